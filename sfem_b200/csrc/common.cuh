// Shared device/host helpers for the b200fem kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>
#include <string>

#include "b200fem.h"

namespace b200fem {

void set_error(const std::string &msg);
bool cuda_ok(cudaError_t e, const char *what, const char *file, int line);

#define B2_CUDA(call)                                                          \
    do {                                                                       \
        if (!::b200fem::cuda_ok((call), #call, __FILE__, __LINE__)) return B200FEM_FAILURE; \
    } while (0)

#define B2_FAIL(msg)                       \
    do {                                   \
        ::b200fem::set_error(msg);         \
        return B200FEM_FAILURE;            \
    } while (0)

#define B2_CHECK_LAUNCH() B2_CUDA(cudaGetLastError())

static inline cudaStream_t as_stream(void *s) { return reinterpret_cast<cudaStream_t>(s); }

int sm_count();  // SM count of the current device (per-device cache, thread-safe)

// The dynamic shared-memory opt-in (cudaFuncAttributeMaxDynamicSharedMemorySize) is a per-device attribute of a kernel:
// one record per launch site remembers, per device, the largest size already granted. Safe across host threads (a
// racing duplicate cudaFuncSetAttribute is harmless).
struct SmemOptIn {
    static constexpr int kMaxDevices = 64;
    std::atomic<size_t> granted[kMaxDevices];
    SmemOptIn() {
        for (auto &g : granted) g.store(0);
    }
    template <typename K>
    int ensure(K kernel, size_t smem) {
        int dev = 0;
        B2_CUDA(cudaGetDevice(&dev));
        if (dev < 0 || dev >= kMaxDevices) dev = kMaxDevices - 1, granted[dev].store(0);  // no caching beyond the table
        if (smem > granted[dev].load(std::memory_order_acquire)) {
            B2_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            granted[dev].store(smem, std::memory_order_release);
        }
        return B200FEM_SUCCESS;
    }
};

// fire-and-forget global add: compiles to RED.E.ADD.F64/F32 (no return value requested)
__device__ __forceinline__ void red_add(double *addr, double v) { atomicAdd(addr, v); }
__device__ __forceinline__ void red_add(float *addr, float v) { atomicAdd(addr, v); }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <typename T>
__device__ __forceinline__ T ldg_stream(const T *p) {
    return __ldcs(p);  // evict-first: connectivity / geometry are read exactly once per apply
}

// Owning device buffer (plan storage)
template <typename T>
struct DeviceArray {
    T *ptr = nullptr;
    size_t n = 0;
    DeviceArray() = default;
    DeviceArray(const DeviceArray &) = delete;
    DeviceArray &operator=(const DeviceArray &) = delete;
    ~DeviceArray() { release(); }
    void release() {
        if (ptr) cudaFree(ptr);
        ptr = nullptr;
        n = 0;
    }
    int alloc(size_t count) {
        release();
        n = count;
        if (count == 0) return B200FEM_SUCCESS;
        B2_CUDA(cudaMalloc((void **)&ptr, count * sizeof(T)));
        return B200FEM_SUCCESS;
    }
    int upload(const T *host, size_t count) {
        if (alloc(count) != B200FEM_SUCCESS) return B200FEM_FAILURE;
        if (count) B2_CUDA(cudaMemcpy(ptr, host, count * sizeof(T), cudaMemcpyHostToDevice));
        return B200FEM_SUCCESS;
    }
    size_t bytes() const { return n * sizeof(T); }
};

// Operator plan base: everything behind b200fem_op_t
struct OpBase {
    int elem_type = 0;
    int level = 0;
    int real_type = B200FEM_F64;
    int block_size = 1;
    ptrdiff_t ne = 0, nn = 0;
    int nxe = 0;
    double geom_bytes_per_elem = 0;
    int scatter = B200FEM_SCATTER_AUTO;
    virtual ~OpBase() {}
    virtual int apply(const void *x, void *y, bool overwrite, cudaStream_t s) = 0;
    virtual int apply_range(const void *x, void *y, ptrdiff_t e0, ptrdiff_t e1, int mode, cudaStream_t s) {
        (void)x; (void)y; (void)e0; (void)e1; (void)mode; (void)s;
        set_error("apply_range not implemented for this operator");
        return B200FEM_FAILURE;
    }
    virtual int apply_overwrite_after(const void *x, void *y, ptrdiff_t e_dep, const unsigned long long *flag, unsigned long long want,
                                      cudaStream_t s) {
        (void)x; (void)y; (void)e_dep; (void)flag; (void)want; (void)s;
        set_error("apply_overwrite_after (in-kernel ghost wait) not implemented for this operator");
        return B200FEM_FAILURE;
    }
    virtual int diag(void *d, cudaStream_t s) {
        (void)d; (void)s;
        set_error("diag not implemented for this operator");
        return B200FEM_FAILURE;
    }
    virtual int block_diag_sym(void *d, cudaStream_t s) {
        (void)d; (void)s;
        set_error("block_diag_sym not implemented for this operator");
        return B200FEM_FAILURE;
    }
    virtual double streamed_bytes(bool overwrite) const { (void)overwrite; return algorithmic_bytes(); }
    virtual int launches(bool overwrite) const { (void)overwrite; return 1; }
    virtual int set_option(const char *key, int value);
    virtual int get_option(const char *key, int *value);
    double algorithmic_bytes() const {
        return double(ne) * (double(nxe) * 4.0 + geom_bytes_per_elem) + double(nn) * block_size * 2.0 * real_type;
    }
};

// factories implemented in the per-element translation units
OpBase *make_tet4_laplacian(ptrdiff_t ne, ptrdiff_t nn, const int32_t *const *elements, const float *fff, int real_type);
OpBase *make_hex8_laplacian(ptrdiff_t ne, ptrdiff_t nn, const int32_t *const *elements, const float *fff, int real_type);
OpBase *make_sshex8_laplacian(int level, ptrdiff_t ne, ptrdiff_t nn, const int32_t *const *elements, const float *fff,
                              int real_type);
OpBase *make_hex8_linear_elasticity(ptrdiff_t ne, ptrdiff_t nn, const int32_t *const *elements, const float *adjugate,
                                    const float *determinant, double mu, double lambda, int real_type);
OpBase *make_sshex8_linear_elasticity(int level, ptrdiff_t ne, ptrdiff_t nn, const int32_t *const *elements,
                                      const float *adjugate, const float *determinant, double mu, double lambda,
                                      int real_type);

OpBase *make_tet4_linear_elasticity(ptrdiff_t ne, ptrdiff_t nn, const int32_t *const *elements, const float *adjugate,
                                    const float *determinant, double mu, double lambda, int real_type);
OpBase *make_p2tet_laplacian(int element_type, ptrdiff_t ne, ptrdiff_t nn, const int32_t *const *elements, const float *fff, int real_type);
OpBase *make_p2tet_linear_elasticity(int element_type, ptrdiff_t ne, ptrdiff_t nn, const int32_t *const *elements, const float *adjugate,
                                     const float *determinant, double mu, double lambda, int real_type);
int packed_laplacian_set_two_pass(OpBase *op, ptrdiff_t n_rows, const ptrdiff_t *ptr, const ptrdiff_t *idx, const int32_t *dest);
OpBase *make_p2tet_from_points(int element_type, bool elasticity, ptrdiff_t ne, ptrdiff_t nn, const int32_t *const *elements,
                               const float *const *points, double mu, double lambda, int real_type);
OpBase *make_packed_laplacian(int element_type, ptrdiff_t n_packs, ptrdiff_t epp, ptrdiff_t ne, ptrdiff_t nn, ptrdiff_t max_nodes,
                              const uint16_t *const *elements, const float *fff, const ptrdiff_t *owned_nodes_ptr, const ptrdiff_t *n_shared,
                              const ptrdiff_t *ghost_ptr, const int32_t *ghost_idx, int real_type);
OpBase *make_sshex8_vector_laplacian(int level, ptrdiff_t ne, ptrdiff_t nn, const int32_t *const *elements, const float *fff, int vector_size,
                                     int real_type);
OpBase *make_sshex8_lumped_mass(int level, ptrdiff_t ne, ptrdiff_t nn, const int32_t *const *elements, const float *const *points,
                                int real_type);
OpBase *make_tet4_laplacian_from_points(ptrdiff_t ne, ptrdiff_t nn, const int32_t *const *elements, const float *const *points,
                                        int real_type);
OpBase *make_tet4_le_from_points(ptrdiff_t ne, ptrdiff_t nn, const int32_t *const *elements, const float *const *points, double mu,
                                 double lambda, int real_type);
OpBase *make_hex8_vector_laplacian(ptrdiff_t ne, ptrdiff_t nn, const int32_t *const *elements, const float *fff, int vector_size,
                                   int real_type);
OpBase *make_mass_from_points(int element_type, ptrdiff_t ne, ptrdiff_t nn, const int32_t *const *elements, const float *const *points,
                              int real_type);
OpBase *make_hex8_le_isoparametric(ptrdiff_t ne, ptrdiff_t nn, const int32_t *const *elements, const float *const *points, double mu,
                                   double lambda, int real_type);
OpBase *make_sshex8_nonaffine_laplacian(int level, ptrdiff_t ne, ptrdiff_t nn, const int32_t *const *elements,
                                        const float *const *points, int real_type);
OpBase *make_sshex8_em(int level, ptrdiff_t ne, ptrdiff_t nn, const int32_t *const *elements, int block_size,
                       const void *element_matrix, int real_type);
int sshex8_laplacian_element_matrix_build(int level, ptrdiff_t ne, const float *fff, int real_type, void *values_host);
int sshex8_le_element_matrix_build(int level, ptrdiff_t ne, const float *adjugate, const float *determinant, double mu, double lambda,
                                   int real_type, void *values_host);

// Greedy element colouring on the host (nodes shared => different colour). Returns the element
// permutation grouped by colour and the colour offsets.
void color_elements(ptrdiff_t ne, ptrdiff_t nn, int nxe, const int32_t *const *elements, std::string *err,
                    int32_t **perm_out, int32_t **offsets_out, int *ncolors_out);

}  // namespace b200fem
