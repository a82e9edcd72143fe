#pragma once
namespace smesh {
    class Communicator {};
    class Context {};
}  // namespace smesh
