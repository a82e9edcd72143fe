/* Glue for the out-of-tree build of the reference operator sources (test infrastructure only).
 * base/sfem_base.hpp:78 declares sfem_abort(); its definition lives in base/sfem_base.cpp, which
 * pulls in MPI. The kernels only reach it through SFEM_ERROR on unrecoverable input. */
#include <cstdio>
#include <cstdlib>
extern "C" void sfem_abort() {
    std::fprintf(stderr, "[oracle/_ref] sfem_abort()\n");
    std::abort();
}
extern "C" int oracle_ref_sizeof_real() {
#ifdef ORACLE_REAL_FLOAT32
    return 4;
#else
    return 8;
#endif
}
