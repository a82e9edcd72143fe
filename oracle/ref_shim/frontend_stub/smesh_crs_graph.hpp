#pragma once
#include "smesh_forward_declarations.hpp"
namespace smesh {
    template <typename C, typename I>
    class CRSGraph {};
}  // namespace smesh
