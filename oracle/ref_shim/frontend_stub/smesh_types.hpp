/* Names of external/smesh that the reference's frontend/sfem_aliases.hpp re-exports into namespace sfem. Only Buffer and
 * the enums are functional; the helper function templates are declared so that the using-declarations compile. */
#pragma once
#include <cstddef>
#include <cstdlib>
#include <functional>
#include <memory>
#include <string>

#include "sfem_base.hpp" /* sfem_config.h: real_t, idx_t, count_t, geom_t, mask_t */

namespace smesh {
    using count_t = ::count_t;
    using idx_t   = ::idx_t;
    using mask_t  = ::mask_t;
    using real_t  = ::real_t;
    using geom_t  = ::geom_t;

    enum ExecutionSpace { EXECUTION_SPACE_HOST = 0, EXECUTION_SPACE_DEVICE = 1, EXECUTION_SPACE_INVALID = -1 };
    enum MemorySpace { MEMORY_SPACE_HOST = 0, MEMORY_SPACE_DEVICE = 1, MEMORY_SPACE_INVALID = -1 };
    enum PrimitiveType { SMESH_DEFAULT = 0, SMESH_FLOAT32 = 4, SMESH_FLOAT64 = 8 };
    inline ExecutionSpace execution_space_from_string(const std::string &s) {
        return s == "device" ? EXECUTION_SPACE_DEVICE : EXECUTION_SPACE_HOST;
    }

    template <typename T>
    class Buffer {
    public:
        Buffer(size_t n, T *ptr, std::function<void(void *)> deleter, MemorySpace ms) : n_(n), ptr_(ptr), deleter_(std::move(deleter)), ms_(ms) {}
        ~Buffer() {
            if (deleter_ && ptr_) deleter_((void *)ptr_);
        }
        Buffer(const Buffer &)            = delete;
        Buffer &operator=(const Buffer &) = delete;
        static std::shared_ptr<Buffer<T>> own(size_t n, T *ptr, std::function<void(void *)> deleter, MemorySpace ms = MEMORY_SPACE_HOST) {
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
        return Buffer<T>::own(n, (T *)calloc(n ? n : 1, sizeof(T)), [](void *p) { free(p); });
    }
    template <typename T>
    SharedBuffer<T> create_buffer(size_t n, ExecutionSpace) {
        return create_host_buffer<T>(n);
    }
    /* declared only: the names frontend/sfem_aliases.hpp pulls into namespace sfem */
    template <typename T, typename S> SharedBuffer<T> astype(const SharedBuffer<S> &);
    template <typename T> SharedBuffer<T *> convert_host_buffer_to_fake_SoA(const SharedBuffer<T> &, int);
    template <typename T> void copy(const SharedBuffer<T> &, const SharedBuffer<T> &);
    template <typename T> SharedBuffer<T *> create_host_buffer_fake_SoA(int, size_t);
    template <typename T> SharedBuffer<T> manage_host_buffer(size_t, T *);
    template <typename T> SharedBuffer<T> soa_to_aos(const SharedBuffer<T *> &);
    template <typename T> SharedBuffer<T> sub(const SharedBuffer<T> &, size_t, size_t);
    template <typename T> SharedBuffer<T> view(const SharedBuffer<T> &, size_t, size_t);
    template <typename T> SharedBuffer<T> zeros_like(const SharedBuffer<T> &);
}  // namespace smesh
