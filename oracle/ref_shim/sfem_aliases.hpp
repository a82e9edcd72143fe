/* Stand-in for frontend/sfem_aliases.hpp of the reference, which only re-exports names from
 * external/smesh (submodule absent from the checkout). TEST INFRASTRUCTURE ONLY.
 * Provides exactly what the header-only solvers under algebra/ use: Buffer<T> (own / wrap / data / size /
 * mem_space), SharedBuffer<T>, create_buffer / create_host_buffer, the ExecutionSpace / MemorySpace enums and
 * the tracing macro. Device allocations go through two hooks so that the reference's solver headers can be
 * driven with DEVICE pointers by the drop-in test (oracle/ref_shim/ref_solvers.cpp). */
#ifndef ORACLE_SHIM_SFEM_ALIASES_HPP
#define ORACLE_SHIM_SFEM_ALIASES_HPP

#include <cassert>
#include <cstddef>
#include <cstdlib>
#include <functional>
#include <memory>

#include "sfem_base.hpp"

namespace smesh {
    enum ExecutionSpace { EXECUTION_SPACE_HOST = 0, EXECUTION_SPACE_DEVICE = 1, EXECUTION_SPACE_INVALID = -1 };
    enum MemorySpace { MEMORY_SPACE_HOST = 0, MEMORY_SPACE_DEVICE = 1, MEMORY_SPACE_INVALID = -1 };

    /* hooks installed by ref_solvers.cpp when a device BLAS table is registered */
    extern "C" void *(*oracle_shim_device_calloc)(size_t bytes);
    extern "C" void (*oracle_shim_device_free)(void *ptr);

    template <typename T>
    class Buffer {
    public:
        Buffer(size_t n, T *ptr, std::function<void(void *)> deleter, MemorySpace ms)
            : n_(n), ptr_(ptr), deleter_(std::move(deleter)), ms_(ms) {}
        ~Buffer() {
            if (deleter_ && ptr_) deleter_((void *)ptr_);
        }
        Buffer(const Buffer &)            = delete;
        Buffer &operator=(const Buffer &) = delete;

        static std::shared_ptr<Buffer<T>> own(size_t n, T *ptr, std::function<void(void *)> deleter,
                                              MemorySpace ms = MEMORY_SPACE_HOST) {
            return std::make_shared<Buffer<T>>(n, ptr, std::move(deleter), ms);
        }
        static std::shared_ptr<Buffer<T>> wrap(size_t n, T *ptr, MemorySpace ms = MEMORY_SPACE_HOST) {
            return std::make_shared<Buffer<T>>(n, ptr, nullptr, ms);
        }
        T          *data() { return ptr_; }
        const T    *data() const { return ptr_; }
        size_t      size() const { return n_; }
        MemorySpace mem_space() const { return ms_; }

    private:
        size_t                      n_;
        T                          *ptr_;
        std::function<void(void *)> deleter_;
        MemorySpace                 ms_;
    };

    template <typename T>
    using SharedBuffer = std::shared_ptr<Buffer<T>>;

    template <typename T>
    SharedBuffer<T> create_host_buffer(size_t n) {
        return Buffer<T>::own(n, (T *)calloc(n ? n : 1, sizeof(T)), [](void *p) { free(p); }, MEMORY_SPACE_HOST);
    }

    template <typename T>
    SharedBuffer<T> create_buffer(size_t n, ExecutionSpace es) {
        if (es == EXECUTION_SPACE_DEVICE) {
            assert(oracle_shim_device_calloc);
            return Buffer<T>::own(
                    n, (T *)oracle_shim_device_calloc((n ? n : 1) * sizeof(T)),
                    [](void *p) { oracle_shim_device_free(p); }, MEMORY_SPACE_DEVICE);
        }
        return create_host_buffer<T>(n);
    }
}  // namespace smesh

namespace sfem {
    using count_t = ::count_t;
    using idx_t   = ::idx_t;
    using mask_t  = ::mask_t;
    using real_t  = ::real_t;

    template <typename T>
    using Buffer = smesh::Buffer<T>;
    template <typename T>
    using SharedBuffer = smesh::SharedBuffer<T>;

    using smesh::create_buffer;
    using smesh::create_host_buffer;

    using smesh::EXECUTION_SPACE_DEVICE;
    using smesh::EXECUTION_SPACE_HOST;
    using smesh::EXECUTION_SPACE_INVALID;
    using smesh::ExecutionSpace;
    using smesh::MEMORY_SPACE_DEVICE;
    using smesh::MEMORY_SPACE_HOST;
    using smesh::MEMORY_SPACE_INVALID;
    using smesh::MemorySpace;
}  // namespace sfem

#define SFEM_TRACE_SCOPE(name)
#define SFEM_TRACE_SCOPE_VARIANT(format, num)
#endif
