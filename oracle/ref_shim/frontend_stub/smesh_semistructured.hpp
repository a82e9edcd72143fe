#pragma once
#include "smesh_mesh.hpp"
namespace smesh {
    inline int semistructured_level(const Mesh &m) { return m.level_; }
}  // namespace smesh
