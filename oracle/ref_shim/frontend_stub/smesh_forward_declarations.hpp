#pragma once
#include <memory>
namespace smesh {
    class Mesh;
    template <typename T>
    class Buffer;
    class Communicator;
    class Context;
    class Sideset;
    template <typename C, typename I>
    class CRSGraph;
    template <typename pack_idx_t>
    class PackedMesh;
}  // namespace smesh
