#pragma once
#include "smesh_forward_declarations.hpp"
namespace smesh {
    template <typename pack_idx_t>
    class PackedMesh {};
}  // namespace smesh
