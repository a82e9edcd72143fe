// Runtime helpers and the operator-plan half of the C ABI (include/b200fem.h).
#include <string.h>

#include <algorithm>
#include <vector>

#include "common.cuh"

namespace b200fem {

static thread_local std::string g_last_error;

void set_error(const std::string &msg) { g_last_error = msg; }

bool cuda_ok(cudaError_t e, const char *what, const char *file, int line) {
    if (e == cudaSuccess) return true;
    char buf[1024];
    snprintf(buf, sizeof(buf), "CUDA error '%s' (%d) in %s at %s:%d", cudaGetErrorString(e), (int)e, what, file, line);
    g_last_error = buf;
    return false;
}

int sm_count() {
    static std::atomic<int> cached[64];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev >= 0 && dev < 64) {
        const int c = cached[dev].load(std::memory_order_relaxed);
        if (c > 0) return c;
    }
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    if (dev >= 0 && dev < 64) cached[dev].store(n, std::memory_order_relaxed);
    return n;
}

int OpBase::set_option(const char *key, int value) {
    if (!strcmp(key, "scatter")) {
        scatter = value;
        return B200FEM_SUCCESS;
    }
    set_error(std::string("unknown option: ") + key);
    return B200FEM_FAILURE;
}

int OpBase::get_option(const char *key, int *value) {
    if (!strcmp(key, "scatter")) {
        *value = scatter;
        return B200FEM_SUCCESS;
    }
    set_error(std::string("unknown option: ") + key);
    return B200FEM_FAILURE;
}

// Greedy colouring, elements visited in their given order; colour = smallest one not used by any
// element sharing a node. 64-bit masks per node (<= 64 colours; hex/tet meshes need 8..~40).
void color_elements(ptrdiff_t ne, ptrdiff_t nn, int nxe, const int32_t *const *elements, std::string *err,
                    int32_t **perm_out, int32_t **offsets_out, int *ncolors_out) {
    std::vector<uint64_t> node_mask((size_t)nn, 0);
    std::vector<int8_t> color((size_t)ne, 0);
    int ncolors = 0;
    for (ptrdiff_t e = 0; e < ne; ++e) {
        uint64_t used = 0;
        for (int v = 0; v < nxe; ++v) used |= node_mask[elements[v][e]];
        int c = 0;
        while (c < 64 && ((used >> c) & 1ull)) ++c;
        if (c >= 64) {
            *err = "element colouring needs more than 64 colours";
            *perm_out = nullptr;
            *offsets_out = nullptr;
            *ncolors_out = 0;
            return;
        }
        color[e] = (int8_t)c;
        ncolors = std::max(ncolors, c + 1);
        for (int v = 0; v < nxe; ++v) node_mask[elements[v][e]] |= (1ull << c);
    }
    int32_t *offsets = (int32_t *)malloc(sizeof(int32_t) * (ncolors + 1));
    int32_t *perm = (int32_t *)malloc(sizeof(int32_t) * (size_t)std::max<ptrdiff_t>(ne, 1));
    std::vector<int32_t> count(ncolors + 1, 0);
    for (ptrdiff_t e = 0; e < ne; ++e) count[color[e] + 1]++;
    for (int c = 0; c < ncolors; ++c) count[c + 1] += count[c];
    for (int c = 0; c <= ncolors; ++c) offsets[c] = count[c];
    for (ptrdiff_t e = 0; e < ne; ++e) perm[count[color[e]]++] = (int32_t)e;
    *perm_out = perm;
    *offsets_out = offsets;
    *ncolors_out = ncolors;
}

}  // namespace b200fem

using namespace b200fem;

struct b200fem_op_s {
    OpBase *impl;
};

extern "C" {

const char *b200fem_last_error(void) { return g_last_error.c_str(); }

int b200fem_device_count(int *count) {
    B2_CUDA(cudaGetDeviceCount(count));
    return B200FEM_SUCCESS;
}

int b200fem_set_device(int device) {
    B2_CUDA(cudaSetDevice(device));
    return B200FEM_SUCCESS;
}

int b200fem_device_info(int *sms, int *cc_major, int *cc_minor, size_t *global_mem_bytes) {
    int dev = 0;
    B2_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp p;
    B2_CUDA(cudaGetDeviceProperties(&p, dev));
    if (sms) *sms = p.multiProcessorCount;
    if (cc_major) *cc_major = p.major;
    if (cc_minor) *cc_minor = p.minor;
    if (global_mem_bytes) *global_mem_bytes = p.totalGlobalMem;
    return B200FEM_SUCCESS;
}

int b200fem_malloc(void **ptr, size_t bytes) {
    *ptr = nullptr;
    if (bytes == 0) return B200FEM_SUCCESS;
    B2_CUDA(cudaMalloc(ptr, bytes));
    B2_CUDA(cudaMemset(*ptr, 0, bytes));
    return B200FEM_SUCCESS;
}

int b200fem_free(void *ptr) {
    if (ptr) B2_CUDA(cudaFree(ptr));
    return B200FEM_SUCCESS;
}

int b200fem_host_malloc(void **ptr, size_t bytes) {
    *ptr = nullptr;
    if (bytes == 0) return B200FEM_SUCCESS;
    B2_CUDA(cudaMallocHost(ptr, bytes));
    return B200FEM_SUCCESS;
}

int b200fem_host_free(void *ptr) {
    if (ptr) B2_CUDA(cudaFreeHost(ptr));
    return B200FEM_SUCCESS;
}

int b200fem_memcpy_h2d(void *dst, const void *src, size_t bytes, void *stream) {
    B2_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, as_stream(stream)));
    return B200FEM_SUCCESS;
}

int b200fem_memcpy_d2h(void *dst, const void *src, size_t bytes, void *stream) {
    B2_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, as_stream(stream)));
    B2_CUDA(cudaStreamSynchronize(as_stream(stream)));
    return B200FEM_SUCCESS;
}

int b200fem_memcpy_d2d(void *dst, const void *src, size_t bytes, void *stream) {
    B2_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, as_stream(stream)));
    return B200FEM_SUCCESS;
}

int b200fem_stream_create(void **stream) {
    cudaStream_t s;
    B2_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    *stream = (void *)s;
    return B200FEM_SUCCESS;
}

int b200fem_stream_destroy(void *stream) {
    B2_CUDA(cudaStreamDestroy(as_stream(stream)));
    return B200FEM_SUCCESS;
}

int b200fem_stream_synchronize(void *stream) {
    B2_CUDA(cudaStreamSynchronize(as_stream(stream)));
    return B200FEM_SUCCESS;
}

int b200fem_device_synchronize(void) {
    B2_CUDA(cudaDeviceSynchronize());
    return B200FEM_SUCCESS;
}

static int check_common(b200fem_op_t *op, ptrdiff_t ne, ptrdiff_t nn, const int32_t *const *elements, int real_type) {
    if (!op) B2_FAIL("op output pointer is NULL");
    *op = nullptr;
    if (ne < 0 || nn < 0) B2_FAIL("negative sizes");
    if (ne > 0 && !elements) B2_FAIL("elements is NULL");
    if (real_type != B200FEM_F32 && real_type != B200FEM_F64) B2_FAIL("real_type must be B200FEM_F32 or B200FEM_F64");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) B2_FAIL("no CUDA device: b200fem has no CPU fallback");
    return B200FEM_SUCCESS;
}

int b200fem_laplacian_create(b200fem_op_t *op, int element_type, int level, ptrdiff_t ne, ptrdiff_t nn,
                             const int32_t *const *elements, const float *fff, int real_type) {
    if (check_common(op, ne, nn, elements, real_type)) return B200FEM_FAILURE;
    if (ne > 0 && !fff) B2_FAIL("fff is NULL");
    OpBase *impl = nullptr;
    switch (element_type) {
        case B200FEM_TET4: impl = make_tet4_laplacian(ne, nn, elements, fff, real_type); break;
        case B200FEM_HEX8: impl = make_hex8_laplacian(ne, nn, elements, fff, real_type); break;
        case B200FEM_TET10:
        case B200FEM_MACRO_TET4: impl = make_p2tet_laplacian(element_type, ne, nn, elements, fff, real_type); break;
        case B200FEM_SSHEX8:
            // apply is built for levels 2, 4 and 8 (the levels of the reference's multigrid hierarchies): refuse the others at
            // creation instead of at the first apply, like the linear-elasticity factory does
            if (level != 2 && level != 4 && level != 8) B2_FAIL("sshex8 laplacian: level must be 2, 4 or 8");
            impl = make_sshex8_laplacian(level, ne, nn, elements, fff, real_type);
            break;
        default: B2_FAIL("laplacian: unsupported element type");
    }
    if (!impl) return B200FEM_FAILURE;
    *op = new b200fem_op_s{impl};
    return B200FEM_SUCCESS;
}

int b200fem_linear_elasticity_create(b200fem_op_t *op, int element_type, int level, ptrdiff_t ne, ptrdiff_t nn,
                                     const int32_t *const *elements, const float *adjugate, const float *determinant,
                                     double mu, double lambda, int real_type) {
    if (check_common(op, ne, nn, elements, real_type)) return B200FEM_FAILURE;
    if (ne > 0 && (!adjugate || !determinant)) B2_FAIL("adjugate/determinant is NULL");
    OpBase *impl = nullptr;
    switch (element_type) {
        case B200FEM_TET4: impl = make_tet4_linear_elasticity(ne, nn, elements, adjugate, determinant, mu, lambda, real_type); break;
        case B200FEM_HEX8: impl = make_hex8_linear_elasticity(ne, nn, elements, adjugate, determinant, mu, lambda, real_type); break;
        case B200FEM_TET10:
        case B200FEM_MACRO_TET4:
            impl = make_p2tet_linear_elasticity(element_type, ne, nn, elements, adjugate, determinant, mu, lambda, real_type);
            break;
        case B200FEM_SSHEX8:
            if (level < 2) B2_FAIL("sshex8 level must be >= 2");
            impl = make_sshex8_linear_elasticity(level, ne, nn, elements, adjugate, determinant, mu, lambda, real_type);
            break;
        default: B2_FAIL("linear elasticity: unsupported element type");
    }
    if (!impl) return B200FEM_FAILURE;
    *op = new b200fem_op_s{impl};
    return B200FEM_SUCCESS;
}

int b200fem_laplacian_create_from_points(b200fem_op_t *op, int element_type, int level, ptrdiff_t ne, ptrdiff_t nn,
                                         const int32_t *const *elements, const float *const *points, int real_type) {
    if (check_common(op, ne, nn, elements, real_type)) return B200FEM_FAILURE;
    if (!points || !points[0] || !points[1] || !points[2]) B2_FAIL("points is NULL");
    OpBase *impl = nullptr;
    switch (element_type) {
        case B200FEM_TET4: impl = make_tet4_laplacian_from_points(ne, nn, elements, points, real_type); break;
        case B200FEM_SSHEX8: impl = make_sshex8_nonaffine_laplacian(level, ne, nn, elements, points, real_type); break;
        case B200FEM_TET10:
        case B200FEM_MACRO_TET4: impl = make_p2tet_from_points(element_type, false, ne, nn, elements, points, 0, 0, real_type); break;
        default: B2_FAIL("laplacian from points: tet4, tet10, macro_tet4 (*_laplacian_apply) and sshex8 (sshex8_laplacian_apply) only");
    }
    if (!impl) return B200FEM_FAILURE;
    *op = new b200fem_op_s{impl};
    return B200FEM_SUCCESS;
}

int b200fem_packed_laplacian_create(b200fem_op_t *op, int element_type, ptrdiff_t n_packs, ptrdiff_t n_elements_per_pack, ptrdiff_t ne,
                                    ptrdiff_t nn, ptrdiff_t max_nodes_per_pack, const uint16_t *const *elements, const float *fff,
                                    const ptrdiff_t *owned_nodes_ptr, const ptrdiff_t *n_shared, const ptrdiff_t *ghost_ptr,
                                    const int32_t *ghost_idx, int real_type) {
    if (!op) B2_FAIL("op output pointer is NULL");
    *op = nullptr;
    if (ne < 0 || nn < 0 || n_packs < 0) B2_FAIL("negative sizes");
    if (real_type != B200FEM_F32 && real_type != B200FEM_F64) B2_FAIL("real_type must be B200FEM_F32 or B200FEM_F64");
    if (!owned_nodes_ptr || !ghost_ptr || (n_packs > 0 && !n_shared)) B2_FAIL("packed laplacian: NULL pack arrays");
    if (ne > 0 && (!elements || !fff)) B2_FAIL("packed laplacian: elements / fff is NULL");
    if (n_packs > 0 && ghost_ptr[n_packs] > 0 && !ghost_idx) B2_FAIL("packed laplacian: ghost_idx is NULL");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) B2_FAIL("no CUDA device: b200fem has no CPU fallback");
    OpBase *impl = make_packed_laplacian(element_type, n_packs, n_elements_per_pack, ne, nn, max_nodes_per_pack, elements, fff,
                                         owned_nodes_ptr, n_shared, ghost_ptr, ghost_idx, real_type);
    if (!impl) return B200FEM_FAILURE;
    *op = new b200fem_op_s{impl};
    return B200FEM_SUCCESS;
}

int b200fem_packed_laplacian_set_two_pass(b200fem_op_t op, ptrdiff_t n_rows, const ptrdiff_t *ghost_reduce_ptr,
                                          const ptrdiff_t *ghost_reduce_idx, const int32_t *ghost_reduce_dest) {
    if (!op || !op->impl) B2_FAIL("invalid operator handle");
    return packed_laplacian_set_two_pass(op->impl, n_rows, ghost_reduce_ptr, ghost_reduce_idx, ghost_reduce_dest);
}

int b200fem_vector_laplacian_create(b200fem_op_t *op, int element_type, ptrdiff_t ne, ptrdiff_t nn, const int32_t *const *elements,
                                    const float *fff, int vector_size, int real_type) {
    if (check_common(op, ne, nn, elements, real_type)) return B200FEM_FAILURE;
    if (ne > 0 && !fff) B2_FAIL("fff is NULL");
    if (vector_size < 1 || vector_size > 16) B2_FAIL("vector laplacian: vector_size out of range");
    if (element_type != B200FEM_HEX8) B2_FAIL("vector laplacian: hex8 only (affine_hex8_vector_laplacian_apply_fff)");
    OpBase *impl = make_hex8_vector_laplacian(ne, nn, elements, fff, vector_size, real_type);
    if (!impl) return B200FEM_FAILURE;
    *op = new b200fem_op_s{impl};
    return B200FEM_SUCCESS;
}

int b200fem_sshex8_vector_laplacian_create(b200fem_op_t *op, int level, ptrdiff_t ne, ptrdiff_t nn, const int32_t *const *elements,
                                           const float *fff, int vector_size, int real_type) {
    if (check_common(op, ne, nn, elements, real_type)) return B200FEM_FAILURE;
    if (ne > 0 && !fff) B2_FAIL("fff is NULL");
    if (vector_size < 1 || vector_size > 16) B2_FAIL("vector laplacian: vector_size out of range");
    if (level != 2 && level != 4 && level != 8) B2_FAIL("sshex8 vector laplacian: level must be 2, 4 or 8");
    OpBase *impl = make_sshex8_vector_laplacian(level, ne, nn, elements, fff, vector_size, real_type);
    if (!impl) return B200FEM_FAILURE;
    *op = new b200fem_op_s{impl};
    return B200FEM_SUCCESS;
}

int b200fem_sshex8_lumped_mass_create(b200fem_op_t *op, int level, ptrdiff_t ne, ptrdiff_t nn, const int32_t *const *elements,
                                      const float *const *points, int real_type) {
    if (check_common(op, ne, nn, elements, real_type)) return B200FEM_FAILURE;
    if (!points || !points[0] || !points[1] || !points[2]) B2_FAIL("points is NULL");
    if (level < 1) B2_FAIL("sshex8 lumped mass: level must be >= 1");
    OpBase *impl = make_sshex8_lumped_mass(level, ne, nn, elements, points, real_type);
    if (!impl) return B200FEM_FAILURE;
    *op = new b200fem_op_s{impl};
    return B200FEM_SUCCESS;
}

int b200fem_mass_create(b200fem_op_t *op, int element_type, ptrdiff_t ne, ptrdiff_t nn, const int32_t *const *elements,
                        const float *const *points, int real_type) {
    if (check_common(op, ne, nn, elements, real_type)) return B200FEM_FAILURE;
    if (!points || !points[0] || !points[1] || !points[2]) B2_FAIL("points is NULL");
    if (element_type != B200FEM_HEX8 && element_type != B200FEM_TET4 && element_type != B200FEM_TET10)
        B2_FAIL("mass: hex8 (apply + lumped), tet4 and tet10 (lumped) only");
    OpBase *impl = make_mass_from_points(element_type, ne, nn, elements, points, real_type);
    if (!impl) return B200FEM_FAILURE;
    *op = new b200fem_op_s{impl};
    return B200FEM_SUCCESS;
}

int b200fem_linear_elasticity_create_from_points(b200fem_op_t *op, int element_type, int level, ptrdiff_t ne, ptrdiff_t nn,
                                                 const int32_t *const *elements, const float *const *points, double mu,
                                                 double lambda, int real_type) {
    (void)level;
    if (check_common(op, ne, nn, elements, real_type)) return B200FEM_FAILURE;
    if (!points || !points[0] || !points[1] || !points[2]) B2_FAIL("points is NULL");
    const bool p2 = element_type == B200FEM_TET10 || element_type == B200FEM_MACRO_TET4;
    if (element_type != B200FEM_HEX8 && element_type != B200FEM_TET4 && !p2)
        B2_FAIL("linear elasticity from points: tet4 / tet10 / macro_tet4 (*_linear_elasticity_apply) and hex8 (hex8_linear_elasticity_apply, isoparametric) only");
    OpBase *impl = p2 ? make_p2tet_from_points(element_type, true, ne, nn, elements, points, mu, lambda, real_type)
                   : element_type == B200FEM_TET4 ? make_tet4_le_from_points(ne, nn, elements, points, mu, lambda, real_type)
                                                  : make_hex8_le_isoparametric(ne, nn, elements, points, mu, lambda, real_type);
    if (!impl) return B200FEM_FAILURE;
    *op = new b200fem_op_s{impl};
    return B200FEM_SUCCESS;
}

int b200fem_sshex8_em_create(b200fem_op_t *op, int level, ptrdiff_t ne, ptrdiff_t nn, const int32_t *const *elements,
                             int block_size, const void *element_matrix, int real_type) {
    if (check_common(op, ne, nn, elements, real_type)) return B200FEM_FAILURE;
    if (ne > 0 && !element_matrix) B2_FAIL("element_matrix is NULL");
    OpBase *impl = make_sshex8_em(level, ne, nn, elements, block_size, element_matrix, real_type);
    if (!impl) return B200FEM_FAILURE;
    *op = new b200fem_op_s{impl};
    return B200FEM_SUCCESS;
}

int b200fem_sshex8_laplacian_element_matrix(int level, ptrdiff_t ne, const float *fff, int real_type, void *values_host) {
    if (real_type != B200FEM_F32 && real_type != B200FEM_F64) B2_FAIL("real_type must be B200FEM_F32 or B200FEM_F64");
    if (level < 1 || ne < 0 || (ne > 0 && (!fff || !values_host))) B2_FAIL("sshex8_laplacian_element_matrix: bad arguments");
    return sshex8_laplacian_element_matrix_build(level, ne, fff, real_type, values_host);
}

int b200fem_sshex8_linear_elasticity_element_matrix(int level, ptrdiff_t ne, const float *adjugate, const float *determinant,
                                                    double mu, double lambda, int real_type, void *values_host) {
    if (real_type != B200FEM_F32 && real_type != B200FEM_F64) B2_FAIL("real_type must be B200FEM_F32 or B200FEM_F64");
    if (level < 1 || ne < 0 || (ne > 0 && (!adjugate || !determinant || !values_host)))
        B2_FAIL("sshex8_linear_elasticity_element_matrix: bad arguments");
    return sshex8_le_element_matrix_build(level, ne, adjugate, determinant, mu, lambda, real_type, values_host);
}

int b200fem_op_destroy(b200fem_op_t op) {
    if (op) {
        delete op->impl;
        delete op;
    }
    return B200FEM_SUCCESS;
}

#define B2_OP_CHECK(op)                      \
    if (!(op) || !(op)->impl) B2_FAIL("invalid operator handle")

int b200fem_op_apply(b200fem_op_t op, const void *x, void *y, void *stream) {
    B2_OP_CHECK(op);
    if (op->impl->ne == 0) return B200FEM_SUCCESS;
    if (!x || !y) B2_FAIL("apply: NULL vector");
    return op->impl->apply(x, y, false, as_stream(stream));
}

int b200fem_op_apply_overwrite(b200fem_op_t op, const void *x, void *y, void *stream) {
    B2_OP_CHECK(op);
    if (!y) B2_FAIL("apply: NULL vector");
    if (op->impl->ne == 0) {
        B2_CUDA(cudaMemsetAsync(y, 0, (size_t)op->impl->nn * op->impl->block_size * op->impl->real_type, as_stream(stream)));
        return B200FEM_SUCCESS;
    }
    if (!x) B2_FAIL("apply: NULL vector");
    return op->impl->apply(x, y, true, as_stream(stream));
}

int b200fem_op_apply_range(b200fem_op_t op, const void *x, void *y, ptrdiff_t e_begin, ptrdiff_t e_end, int mode,
                           void *stream) {
    B2_OP_CHECK(op);
    if (!x || !y) B2_FAIL("apply_range: NULL vector");
    if (e_begin < 0 || e_end < e_begin || e_end > op->impl->ne) B2_FAIL("apply_range: element range out of bounds");
    if (mode < 0 || mode > 2) B2_FAIL("apply_range: mode must be 0, 1 or 2");
    return op->impl->apply_range(x, y, e_begin, e_end, mode, as_stream(stream));
}

int b200fem_op_diag(b200fem_op_t op, void *diag, void *stream) {
    B2_OP_CHECK(op);
    if (op->impl->ne == 0) return B200FEM_SUCCESS;
    if (!diag) B2_FAIL("diag: NULL vector");
    return op->impl->diag(diag, as_stream(stream));
}

int b200fem_op_block_diag_sym(b200fem_op_t op, void *bdiag, void *stream) {
    B2_OP_CHECK(op);
    if (op->impl->ne == 0) return B200FEM_SUCCESS;
    if (!bdiag) B2_FAIL("block_diag_sym: NULL vector");
    return op->impl->block_diag_sym(bdiag, as_stream(stream));
}

ptrdiff_t b200fem_op_rows(b200fem_op_t op) { return (op && op->impl) ? op->impl->nn * op->impl->block_size : -1; }
int b200fem_op_apply_overwrite_after(b200fem_op_t op, const void *x, void *y, ptrdiff_t e_dependent, const unsigned long long *ready_flag,
                                     unsigned long long ready_value, void *stream) {
    B2_OP_CHECK(op);
    if (!x || !y) B2_FAIL("apply_overwrite_after: NULL vector");
    if (e_dependent < 0 || e_dependent > op->impl->ne) B2_FAIL("apply_overwrite_after: first dependent element out of bounds");
    if (!ready_flag) return op->impl->apply(x, y, true, as_stream(stream));
    return op->impl->apply_overwrite_after(x, y, e_dependent, ready_flag, ready_value, as_stream(stream));
}
ptrdiff_t b200fem_op_nelements(b200fem_op_t op) { return (op && op->impl) ? op->impl->ne : -1; }
int b200fem_op_block_size(b200fem_op_t op) { return (op && op->impl) ? op->impl->block_size : -1; }
int b200fem_op_real_type(b200fem_op_t op) { return (op && op->impl) ? op->impl->real_type : -1; }
double b200fem_op_algorithmic_bytes(b200fem_op_t op) { return (op && op->impl) ? op->impl->algorithmic_bytes() : -1.0; }
double b200fem_op_streamed_bytes(b200fem_op_t op, int overwrite) {
    return (op && op->impl) ? op->impl->streamed_bytes(overwrite != 0) : -1.0;
}
int b200fem_op_gpu_launches_per_apply(b200fem_op_t op, int overwrite) {
    return (op && op->impl) ? op->impl->launches(overwrite != 0) : -1;
}
int b200fem_op_set_option(b200fem_op_t op, const char *key, int value) {
    B2_OP_CHECK(op);
    return op->impl->set_option(key, value);
}
int b200fem_op_get_option(b200fem_op_t op, const char *key, int *value) {
    B2_OP_CHECK(op);
    return op->impl->get_option(key, value);
}

}  // extern "C"
