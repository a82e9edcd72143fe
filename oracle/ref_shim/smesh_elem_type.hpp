/* Minimal stand-in for external/smesh's element-type enum (submodule absent from the
 * checkout). Only the names the operator sources reference are provided. */
#ifndef ORACLE_SHIM_SMESH_ELEM_TYPE_HPP
#define ORACLE_SHIM_SMESH_ELEM_TYPE_HPP
namespace smesh {
    enum ElemType {
        INVALID = -1, NODE1 = 1, EDGE2 = 2, TRI3 = 3, TET4 = 4, TRI6 = 6, HEX8 = 8, TET10 = 10,
        QUAD4 = 40, MACRO_TRI3 = 60, MACRO_TET4 = 100, SSHEX8 = 800
    };
    static inline const char *type_to_string(int t) {
        switch (t) { case TET4: return "TET4"; case HEX8: return "HEX8"; case SSHEX8: return "SSHEX8"; default: return "?"; }
    }
    static inline bool is_semistructured_type(int t) { return t == SSHEX8; }
    static inline int semistructured_level(int) { return 0; }
    static inline ElemType macro_base_elem(ElemType t) { return t == SSHEX8 ? HEX8 : (t == MACRO_TET4 ? TET4 : (t == MACRO_TRI3 ? TRI3 : t)); }
    static inline ElemType macro_type_variant(ElemType t) { return t == TET10 ? MACRO_TET4 : (t == TRI6 ? MACRO_TRI3 : t); }
}
#endif
