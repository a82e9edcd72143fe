// 10-node tetrahedra: tet10 (P2) and macro_tet4 (P1-iso-P2) Laplacian and linear elasticity from the
// cached tet4 geometry of the 4 corners, matrix-free:  y[ev] += K_e u[ev].
//
// Replaces
//   tet10_laplacian_apply_opt / _diag_opt                 operators/tet10/tet10_laplacian.cpp:232-302
//   macro_tet4_laplacian_apply_opt / _diag_opt            operators/macro_tet4/macro_tet4_laplacian.cpp:401-490
//   tet10_linear_elasticity_apply_opt / _diag_opt         operators/tet10/tet10_linear_elasticity.cpp:209-355
//   macro_tet4_linear_elasticity_apply_opt / _diag_opt    operators/macro_tet4/macro_tet4_linear_elasticity.cpp:206-277, 397-437
// Local operators: p2_tet_local.cuh (corner-gradient form of P2, table-driven sub-tets).
//
// Data layout in HBM (plan): connectivity SoA int32 [10][ne]; FFF SoA fp32 [6][ne] or adjugate SoA fp32 [9][ne] +
// determinant [ne], streamed once per apply. A 10-node tet shares each corner with ~24 and each edge node with ~5
// elements. Two scatters: B200FEM_SCATTER_PATCH, the patch kernel of patch.cuh (256 consecutive elements staged in shared
// memory, one RED per unique node of the patch; the plan is built on first use), and B200FEM_SCATTER_ATOMIC, the
// thread-per-element form with one RED per (element, node, component). The default per operator is the faster one as
// measured on B200 (profiles/r02p2_extras.jsonl): ATOMIC for both Laplacians and macro_tet4 elasticity, PATCH for tet10
// elasticity.
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "coloring.cuh"
#include "common.cuh"
#include "p2_tet_local.cuh"
#include "patch.cuh"

namespace b200fem {

template <typename T, bool MACRO>
struct P2TetLaplacianLocal {
    const float *fff;  // SoA [6][ne]
    ptrdiff_t ne;
    struct Geo {
        float f[6];
    };
    __device__ __forceinline__ void load(const ptrdiff_t e, Geo &g) const {
#pragma unroll
        for (int k = 0; k < 6; ++k) g.f[k] = ldg_stream(&fff[k * ne + e]);
    }
    __device__ __forceinline__ void compute(const Geo &g, const T (&uin)[10], T (&out)[10]) const {
        T f[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) f[k] = (T)g.f[k];
        if (MACRO)
            macro_tet4_laplacian_local(f, uin, out);
        else
            tet10_laplacian_local(f, uin, out);
    }
};

template <typename T, bool MACRO>
struct P2TetLELocal {
    const float *adjugate;     // SoA [9][ne]
    const float *determinant;  // [ne]
    ptrdiff_t ne;
    T mu, lambda;
    struct Geo {
        float a[9], det;
    };
    __device__ __forceinline__ void load(const ptrdiff_t e, Geo &g) const {
#pragma unroll
        for (int k = 0; k < 9; ++k) g.a[k] = ldg_stream(&adjugate[k * ne + e]);
        g.det = ldg_stream(&determinant[e]);
    }
    __device__ __forceinline__ void compute(const Geo &g, const T (&uin)[30], T (&out)[30]) const {
        T A[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) A[k] = (T)g.a[k];
        T ul[10][3], ol[10][3];
#pragma unroll
        for (int v = 0; v < 10; ++v)
#pragma unroll
            for (int b = 0; b < 3; ++b) ul[v][b] = uin[v * 3 + b];
        p2tet_le_local<T, MACRO>(A, (T)g.det, mu, lambda, ul, ol);
#pragma unroll
        for (int v = 0; v < 10; ++v)
#pragma unroll
            for (int b = 0; b < 3; ++b) out[v * 3 + b] = ol[v][b];
    }
};

// thread-per-element forms (B200FEM_SCATTER_ATOMIC, and the fallback when a patch does not fit shared memory)
template <typename T, bool MACRO>
__global__ void __launch_bounds__(128)
    p2tet_laplacian_apply_kernel(const ptrdiff_t ne, const int32_t *__restrict__ elements, const float *__restrict__ fff,
                                 const T *__restrict__ u, T *__restrict__ y) {
    for (ptrdiff_t e = (ptrdiff_t)blockIdx.x * blockDim.x + threadIdx.x; e < ne; e += (ptrdiff_t)gridDim.x * blockDim.x) {
        int32_t ev[10];
#pragma unroll
        for (int v = 0; v < 10; ++v) ev[v] = ldg_stream(&elements[v * ne + e]);
        T f[6], ul[10], ol[10];
#pragma unroll
        for (int k = 0; k < 6; ++k) f[k] = (T)ldg_stream(&fff[k * ne + e]);
#pragma unroll
        for (int v = 0; v < 10; ++v) ul[v] = u[ev[v]];
        if (MACRO)
            macro_tet4_laplacian_local(f, ul, ol);
        else
            tet10_laplacian_local(f, ul, ol);
#pragma unroll
        for (int v = 0; v < 10; ++v) red_add(&y[ev[v]], ol[v]);
    }
}

template <typename T, bool MACRO>
__global__ void __launch_bounds__(128)
    p2tet_laplacian_diag_kernel(const ptrdiff_t ne, const int32_t *__restrict__ elements, const float *__restrict__ fff,
                                const bool reference_scale, T *__restrict__ diag) {
    for (ptrdiff_t e = (ptrdiff_t)blockIdx.x * blockDim.x + threadIdx.x; e < ne; e += (ptrdiff_t)gridDim.x * blockDim.x) {
        T f[6], ol[10];
#pragma unroll
        for (int k = 0; k < 6; ++k) f[k] = (T)fff[k * ne + e];
        if (MACRO)
            macro_tet4_laplacian_local_diag(f, ol);
        else
            tet10_laplacian_local_diag(f, reference_scale, ol);
#pragma unroll
        for (int v = 0; v < 10; ++v) red_add(&diag[elements[v * ne + e]], ol[v]);
    }
}

template <typename T, bool MACRO>
__global__ void __launch_bounds__(128)
    p2tet_le_apply_kernel(const ptrdiff_t ne, const int32_t *__restrict__ elements, const float *__restrict__ adjugate,
                          const float *__restrict__ determinant, const T mu, const T lambda, const T *__restrict__ u, T *__restrict__ y) {
    for (ptrdiff_t e = (ptrdiff_t)blockIdx.x * blockDim.x + threadIdx.x; e < ne; e += (ptrdiff_t)gridDim.x * blockDim.x) {
        int32_t ev[10];
#pragma unroll
        for (int v = 0; v < 10; ++v) ev[v] = ldg_stream(&elements[v * ne + e]);
        T A[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) A[k] = (T)ldg_stream(&adjugate[k * ne + e]);
        const T det = (T)ldg_stream(&determinant[e]);
        T ul[10][3], ol[10][3];
#pragma unroll
        for (int v = 0; v < 10; ++v)
#pragma unroll
            for (int b = 0; b < 3; ++b) ul[v][b] = u[(ptrdiff_t)ev[v] * 3 + b];
        p2tet_le_local<T, MACRO>(A, det, mu, lambda, ul, ol);
#pragma unroll
        for (int v = 0; v < 10; ++v)
#pragma unroll
            for (int b = 0; b < 3; ++b) red_add(&y[(ptrdiff_t)ev[v] * 3 + b], ol[v][b]);
    }
}

// diagonal: one thread per (element, local dof): row (v, b) of the 30 x 30 element matrix applied to its own unit vector
template <typename T, bool MACRO>
__global__ void __launch_bounds__(128)
    p2tet_le_diag_kernel(const ptrdiff_t ne, const int32_t *__restrict__ elements, const float *__restrict__ adjugate,
                         const float *__restrict__ determinant, const T mu, const T lambda, T *__restrict__ diag) {
    const ptrdiff_t total = ne * 30;
    for (ptrdiff_t i = (ptrdiff_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (ptrdiff_t)gridDim.x * blockDim.x) {
        const ptrdiff_t e = i % ne;  // consecutive threads: consecutive elements of the same local dof
        const int dof = (int)(i / ne), v = dof / 3, b = dof % 3;
        T A[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) A[k] = (T)adjugate[k * ne + e];
        const T det = (T)determinant[e];
        T ul[10][3], ol[10][3];
#pragma unroll
        for (int j = 0; j < 10; ++j)
#pragma unroll
            for (int c = 0; c < 3; ++c) ul[j][c] = (j == v && c == b) ? T(1) : T(0);
        p2tet_le_local<T, MACRO>(A, det, mu, lambda, ul, ol);
        T r = T(0);
#pragma unroll
        for (int j = 0; j < 10; ++j)
#pragma unroll
            for (int c = 0; c < 3; ++c)
                if (j == v && c == b) r = ol[j][c];
        red_add(&diag[(ptrdiff_t)elements[v * ne + e] * 3 + b], r);
    }
}

// From the points (the reference's non-cached paths: tet10_laplacian_apply / _diag, operators/tet10/tet10_laplacian.cpp:65-230;
// macro_tet4_laplacian_apply / _diag, operators/macro_tet4/macro_tet4_laplacian.cpp:127-260; tet10_linear_elasticity_apply / _diag,
// operators/tet10/tet10_linear_elasticity.cpp:21-207; macro_tet4_linear_elasticity_apply / _diag,
// operators/macro_tet4/macro_tet4_linear_elasticity.cpp:137-204, 330-395): the adjugate / determinant / FFF of the corner tet
// recomputed per element in the vector precision (tet4_adjugate_and_det_s / tet4_fff_s), then the same local operators.
template <typename T, bool MACRO, bool LE, bool DIAG>
__global__ void __launch_bounds__(128)
    p2tet_points_kernel(const ptrdiff_t ne, const ptrdiff_t nn, const int32_t *__restrict__ elements, const float *__restrict__ points,
                        const T mu, const T lambda, const bool reference_diag, const T *__restrict__ u, T *__restrict__ y) {
    for (ptrdiff_t e = (ptrdiff_t)blockIdx.x * blockDim.x + threadIdx.x; e < ne; e += (ptrdiff_t)gridDim.x * blockDim.x) {
        int32_t ev[10];
#pragma unroll
        for (int v = 0; v < 10; ++v) ev[v] = ldg_stream(&elements[v * ne + e]);
        // J[r][c] = x_r(corner c + 1) - x_r(corner 0); A = adjugate = det J^-1; FFF = A A^T / (6 det)
        T J[9];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const T p0 = (T)points[(size_t)r * nn + ev[0]];
#pragma unroll
            for (int c = 0; c < 3; ++c) J[r * 3 + c] = (T)points[(size_t)r * nn + ev[c + 1]] - p0;
        }
        T A[9];
        A[0] = J[4] * J[8] - J[5] * J[7];
        A[1] = J[2] * J[7] - J[1] * J[8];
        A[2] = J[1] * J[5] - J[2] * J[4];
        A[3] = J[5] * J[6] - J[3] * J[8];
        A[4] = J[0] * J[8] - J[2] * J[6];
        A[5] = J[2] * J[3] - J[0] * J[5];
        A[6] = J[3] * J[7] - J[4] * J[6];
        A[7] = J[1] * J[6] - J[0] * J[7];
        A[8] = J[0] * J[4] - J[1] * J[3];
        const T det = J[0] * A[0] + J[1] * A[3] + J[2] * A[6];
        if (LE) {
            T ul[10][3], ol[10][3];
            if (DIAG) {
                for (int v = 0; v < 10; ++v)
                    for (int b = 0; b < 3; ++b) {
#pragma unroll
                        for (int j = 0; j < 10; ++j)
#pragma unroll
                            for (int c = 0; c < 3; ++c) ul[j][c] = (j == v && c == b) ? T(1) : T(0);
                        p2tet_le_local<T, MACRO>(A, det, mu, lambda, ul, ol);
                        T r = T(0);
#pragma unroll
                        for (int j = 0; j < 10; ++j)
#pragma unroll
                            for (int c = 0; c < 3; ++c)
                                if (j == v && c == b) r = ol[j][c];
                        red_add(&y[(ptrdiff_t)elements[v * ne + e] * 3 + b], r);
                    }
            } else {
#pragma unroll
                for (int v = 0; v < 10; ++v)
#pragma unroll
                    for (int b = 0; b < 3; ++b) ul[v][b] = u[(ptrdiff_t)ev[v] * 3 + b];
                p2tet_le_local<T, MACRO>(A, det, mu, lambda, ul, ol);
#pragma unroll
                for (int v = 0; v < 10; ++v)
#pragma unroll
                    for (int b = 0; b < 3; ++b) red_add(&y[(ptrdiff_t)ev[v] * 3 + b], ol[v][b]);
            }
        } else {
            const T sc = T(1) / (T(6) * det);
            T f[6];
            int k = 0;
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int j = i; j < 3; ++j) f[k++] = (A[i * 3] * A[j * 3] + A[i * 3 + 1] * A[j * 3 + 1] + A[i * 3 + 2] * A[j * 3 + 2]) * sc;
            T ul[10], ol[10];
            if (DIAG) {
                if (MACRO)
                    macro_tet4_laplacian_local_diag(f, ol);
                else
                    tet10_laplacian_local_diag(f, reference_diag, ol);
            } else {
#pragma unroll
                for (int v = 0; v < 10; ++v) ul[v] = u[ev[v]];
                if (MACRO)
                    macro_tet4_laplacian_local(f, ul, ol);
                else
                    tet10_laplacian_local(f, ul, ol);
            }
#pragma unroll
            for (int v = 0; v < 10; ++v) red_add(&y[ev[v]], ol[v]);
        }
    }
}

static constexpr int kP2PatchE = 256;

// The patch plan is built on first use (PATCH is not the default of every family, see make_p2tet_*): the plan keeps the host
// connectivity for that.
struct P2PatchHolder {
    std::vector<int32_t> host_el;  // [10][ne]
    PatchPlan patch;
    int ensure(ptrdiff_t ne) {
        if (patch.ready) return B200FEM_SUCCESS;
        const int32_t *rows[10];
        for (int v = 0; v < 10; ++v) rows[v] = host_el.data() + (size_t)v * ne;
        return patch.build(10, kP2PatchE, ne, rows);
    }
};

struct P2TetLaplacian final : OpBase {
    bool macro = false;
    DeviceArray<int32_t> elements;  // [10][ne]
    DeviceArray<float> fff;         // [6][ne]
    P2PatchHolder ph;
    PatchPlan &patch = ph.patch;
    int reference_diag = 1;  // tet10: reproduce the reference's diagonal (1/6 of the apply's, see p2_tet_local.cuh)

    template <typename T, bool MACRO>
    int apply_t(const T *x, T *y, cudaStream_t s) {
        if (ne == 0) return B200FEM_SUCCESS;
        if (scatter == B200FEM_SCATTER_PATCH && ph.ensure(ne)) return B200FEM_FAILURE;
        if (scatter == B200FEM_SCATTER_PATCH && patch.ready && patch.smem_bytes_direct(1, sizeof(T)) + (size_t)20 * kP2PatchE + 2 * patch.max_u_pad() <= 200 * 1024)
            return launch_patch_apply_direct<T, 10, 1, kP2PatchE, 256, P2TetLaplacianLocal<T, MACRO>, true, true>(
                patch, P2TetLaplacianLocal<T, MACRO>{fff.ptr, ne}, x, y, s);
        const int grid = (int)std::min<ptrdiff_t>((ne + 127) / 128, (ptrdiff_t)sm_count() * 16);
        p2tet_laplacian_apply_kernel<T, MACRO><<<grid, 128, 0, s>>>(ne, elements.ptr, fff.ptr, x, y);
        B2_CHECK_LAUNCH();
        return B200FEM_SUCCESS;
    }
    int apply(const void *x, void *y, bool overwrite, cudaStream_t s) override {
        if (overwrite) B2_CUDA(cudaMemsetAsync(y, 0, (size_t)nn * real_type, s));
        if (real_type == B200FEM_F64)
            return macro ? apply_t<double, true>((const double *)x, (double *)y, s) : apply_t<double, false>((const double *)x, (double *)y, s);
        return macro ? apply_t<float, true>((const float *)x, (float *)y, s) : apply_t<float, false>((const float *)x, (float *)y, s);
    }
    template <typename T>
    int diag_t(T *d, cudaStream_t s) {
        if (ne == 0) return B200FEM_SUCCESS;
        const int grid = (int)std::min<ptrdiff_t>((ne + 127) / 128, (ptrdiff_t)sm_count() * 16);
        if (macro)
            p2tet_laplacian_diag_kernel<T, true><<<grid, 128, 0, s>>>(ne, elements.ptr, fff.ptr, false, d);
        else
            p2tet_laplacian_diag_kernel<T, false><<<grid, 128, 0, s>>>(ne, elements.ptr, fff.ptr, reference_diag != 0, d);
        B2_CHECK_LAUNCH();
        return B200FEM_SUCCESS;
    }
    int diag(void *d, cudaStream_t s) override { return real_type == B200FEM_F64 ? diag_t((double *)d, s) : diag_t((float *)d, s); }
    int launches(bool overwrite) const override { return 1 + (overwrite ? 1 : 0); }
    double streamed_bytes(bool overwrite) const override {
        const double idx = (scatter == B200FEM_SCATTER_PATCH && patch.ready) ? patch.index_bytes() : 40.0 * ne;
        return idx + 24.0 * ne + double(nn) * real_type * (overwrite ? 2.0 : 3.0);
    }
    int set_option(const char *key, int value) override {
        if (!strcmp(key, "reference_diag")) { reference_diag = value; return B200FEM_SUCCESS; }
        return OpBase::set_option(key, value);
    }
    int get_option(const char *key, int *value) override {
        if (!strcmp(key, "reference_diag")) { *value = reference_diag; return B200FEM_SUCCESS; }
        if (!strcmp(key, "patch_avg_unique_nodes")) { *value = patch.ready ? (int)patch.avg_u : 0; return B200FEM_SUCCESS; }
        return OpBase::get_option(key, value);
    }
};

struct P2TetLinearElasticity final : OpBase {
    bool macro = false;
    DeviceArray<int32_t> elements;  // [10][ne]
    DeviceArray<float> adjugate;    // SoA [9][ne]
    DeviceArray<float> determinant;
    P2PatchHolder ph;
    PatchPlan &patch = ph.patch;
    double mu = 1, lambda = 1;

    template <typename T, bool MACRO>
    int apply_t(const T *x, T *y, cudaStream_t s) {
        if (ne == 0) return B200FEM_SUCCESS;
        if (scatter == B200FEM_SCATTER_PATCH && ph.ensure(ne)) return B200FEM_FAILURE;
        if (scatter == B200FEM_SCATTER_PATCH && patch.ready && patch.smem_bytes_direct(3, sizeof(T)) + (size_t)20 * kP2PatchE + 2 * patch.max_u_pad() <= 200 * 1024)
            return launch_patch_apply_direct<T, 10, 3, kP2PatchE, 256, P2TetLELocal<T, MACRO>, true, false>(
                patch, P2TetLELocal<T, MACRO>{adjugate.ptr, determinant.ptr, ne, (T)mu, (T)lambda}, x, y, s);
        const int grid = (int)std::min<ptrdiff_t>((ne + 127) / 128, (ptrdiff_t)sm_count() * 16);
        p2tet_le_apply_kernel<T, MACRO><<<grid, 128, 0, s>>>(ne, elements.ptr, adjugate.ptr, determinant.ptr, (T)mu, (T)lambda, x, y);
        B2_CHECK_LAUNCH();
        return B200FEM_SUCCESS;
    }
    int apply(const void *x, void *y, bool overwrite, cudaStream_t s) override {
        if (overwrite) B2_CUDA(cudaMemsetAsync(y, 0, (size_t)nn * 3 * real_type, s));
        if (real_type == B200FEM_F64)
            return macro ? apply_t<double, true>((const double *)x, (double *)y, s) : apply_t<double, false>((const double *)x, (double *)y, s);
        return macro ? apply_t<float, true>((const float *)x, (float *)y, s) : apply_t<float, false>((const float *)x, (float *)y, s);
    }
    template <typename T>
    int diag_t(T *d, cudaStream_t s) {
        if (ne == 0) return B200FEM_SUCCESS;
        const int grid = (int)std::min<ptrdiff_t>((ne * 30 + 127) / 128, (ptrdiff_t)sm_count() * 16);
        if (macro)
            p2tet_le_diag_kernel<T, true><<<grid, 128, 0, s>>>(ne, elements.ptr, adjugate.ptr, determinant.ptr, (T)mu, (T)lambda, d);
        else
            p2tet_le_diag_kernel<T, false><<<grid, 128, 0, s>>>(ne, elements.ptr, adjugate.ptr, determinant.ptr, (T)mu, (T)lambda, d);
        B2_CHECK_LAUNCH();
        return B200FEM_SUCCESS;
    }
    int diag(void *d, cudaStream_t s) override { return real_type == B200FEM_F64 ? diag_t((double *)d, s) : diag_t((float *)d, s); }
    int launches(bool overwrite) const override { return 1 + (overwrite ? 1 : 0); }
    double streamed_bytes(bool overwrite) const override {
        const double idx = (scatter == B200FEM_SCATTER_PATCH && patch.ready) ? patch.index_bytes() : 40.0 * ne;
        return idx + 40.0 * ne + double(nn) * 3 * real_type * (overwrite ? 2.0 : 3.0);
    }
    int get_option(const char *key, int *value) override {
        if (!strcmp(key, "patch_avg_unique_nodes")) { *value = patch.ready ? (int)patch.avg_u : 0; return B200FEM_SUCCESS; }
        return OpBase::get_option(key, value);
    }
};

struct P2TetPointsOp final : OpBase {
    bool macro = false, elasticity = false;
    DeviceArray<int32_t> elements;  // [10][ne]
    DeviceArray<float> points;      // [3][nn]
    double mu = 1, lambda = 1;
    int reference_diag = 1;

    template <typename T, bool DIAG>
    int run(const T *x, T *y, cudaStream_t s) {
        if (ne == 0) return B200FEM_SUCCESS;
        const int grid = (int)std::min<ptrdiff_t>((ne + 127) / 128, (ptrdiff_t)sm_count() * 16);
        const bool rd = reference_diag != 0;
#define B2_P2_POINTS(MACRO, LE) \
    p2tet_points_kernel<T, MACRO, LE, DIAG><<<grid, 128, 0, s>>>(ne, nn, elements.ptr, points.ptr, (T)mu, (T)lambda, rd, x, y)
        if (elasticity) {
            if (macro) B2_P2_POINTS(true, true); else B2_P2_POINTS(false, true);
        } else {
            if (macro) B2_P2_POINTS(true, false); else B2_P2_POINTS(false, false);
        }
#undef B2_P2_POINTS
        B2_CHECK_LAUNCH();
        return B200FEM_SUCCESS;
    }
    int apply(const void *x, void *y, bool overwrite, cudaStream_t s) override {
        if (overwrite) B2_CUDA(cudaMemsetAsync(y, 0, (size_t)nn * block_size * real_type, s));
        return real_type == B200FEM_F64 ? run<double, false>((const double *)x, (double *)y, s) : run<float, false>((const float *)x, (float *)y, s);
    }
    int diag(void *d, cudaStream_t s) override {
        return real_type == B200FEM_F64 ? run<double, true>(nullptr, (double *)d, s) : run<float, true>(nullptr, (float *)d, s);
    }
    int launches(bool overwrite) const override { return 1 + (overwrite ? 1 : 0); }
    int set_option(const char *key, int value) override {
        if (!strcmp(key, "reference_diag")) { reference_diag = value; return B200FEM_SUCCESS; }
        return OpBase::set_option(key, value);
    }
    int get_option(const char *key, int *value) override {
        if (!strcmp(key, "reference_diag")) { *value = reference_diag; return B200FEM_SUCCESS; }
        return OpBase::get_option(key, value);
    }
};

static bool upload_connectivity(ptrdiff_t ne, const int32_t *const *elements, std::vector<int32_t> &el, DeviceArray<int32_t> &dst) {
    el.resize((size_t)10 * ne);
    for (int v = 0; v < 10; ++v) std::copy(elements[v], elements[v] + ne, el.begin() + (size_t)v * ne);
    return dst.upload(el.data(), el.size()) != 0;
}

OpBase *make_p2tet_laplacian(int element_type, ptrdiff_t ne, ptrdiff_t nn, const int32_t *const *elements, const float *fff, int real_type) {
    if (!check_indices(10, ne, nn, elements)) {
        set_error("tet10 / macro_tet4 laplacian: node index out of range");
        return nullptr;
    }
    auto *op = new P2TetLaplacian();
    op->macro = element_type == B200FEM_MACRO_TET4;
    op->elem_type = element_type;
    op->real_type = real_type;
    op->ne = ne;
    op->nn = nn;
    op->nxe = 10;
    op->geom_bytes_per_elem = 24;
    // measured on B200 (profiles/r02p2_extras.jsonl, 2.96e6 elements, 4.0e6 nodes, SFC order, fp64): RED per corner 0.153 ms (tet10) /
    // 0.157 ms (macro_tet4) against 0.170 / 0.215 ms for 256-element patches -- a 10-node patch has 650 unique nodes for 256
    // elements (2.5 per element against 1.36 in the mesh), too little reuse to pay for the staging
    op->scatter = B200FEM_SCATTER_ATOMIC;
    std::vector<float> f((size_t)6 * ne);
    for (ptrdiff_t e = 0; e < ne; ++e)
        for (int k = 0; k < 6; ++k) f[(size_t)k * ne + e] = fff[e * 6 + k];
    if (upload_connectivity(ne, elements, op->ph.host_el, op->elements) || op->fff.upload(f.data(), f.size())) {
        delete op;
        return nullptr;
    }
    return op;
}

OpBase *make_p2tet_linear_elasticity(int element_type, ptrdiff_t ne, ptrdiff_t nn, const int32_t *const *elements, const float *adjugate,
                                     const float *determinant, double mu, double lambda, int real_type) {
    if (!check_indices(10, ne, nn, elements)) {
        set_error("tet10 / macro_tet4 linear elasticity: node index out of range");
        return nullptr;
    }
    auto *op = new P2TetLinearElasticity();
    op->macro = element_type == B200FEM_MACRO_TET4;
    op->elem_type = element_type;
    op->real_type = real_type;
    op->block_size = 3;
    op->ne = ne;
    op->nn = nn;
    op->nxe = 10;
    op->geom_bytes_per_elem = 40;
    op->mu = mu;
    op->lambda = lambda;
    // measured (same record): tet10 0.380 ms with patches against 0.510 ms with REDs (30 REDs per element); macro_tet4 0.550 ms with
    // REDs against 0.673 ms with patches (its patch instance needs 244 registers)
    op->scatter = op->macro ? B200FEM_SCATTER_ATOMIC : B200FEM_SCATTER_PATCH;
    std::vector<float> a((size_t)9 * ne);
    for (ptrdiff_t e = 0; e < ne; ++e)
        for (int k = 0; k < 9; ++k) a[(size_t)k * ne + e] = adjugate[e * 9 + k];
    if (upload_connectivity(ne, elements, op->ph.host_el, op->elements) ||
        op->adjugate.upload(a.data(), a.size()) || op->determinant.upload(determinant, (size_t)ne)) {
        delete op;
        return nullptr;
    }
    return op;
}

OpBase *make_p2tet_from_points(int element_type, bool elasticity, ptrdiff_t ne, ptrdiff_t nn, const int32_t *const *elements,
                               const float *const *points, double mu, double lambda, int real_type) {
    if (!check_indices(10, ne, nn, elements)) {
        set_error("tet10 / macro_tet4 (from points): node index out of range");
        return nullptr;
    }
    auto *op = new P2TetPointsOp();
    op->macro = element_type == B200FEM_MACRO_TET4;
    op->elasticity = elasticity;
    op->elem_type = element_type;
    op->real_type = real_type;
    op->block_size = elasticity ? 3 : 1;
    op->ne = ne;
    op->nn = nn;
    op->nxe = 10;
    op->geom_bytes_per_elem = 48;  // the 4 corner coordinates an element reads
    op->mu = mu;
    op->lambda = lambda;
    op->scatter = B200FEM_SCATTER_ATOMIC;
    std::vector<int32_t> el;
    bool bad = upload_connectivity(ne, elements, el, op->elements) || op->points.alloc((size_t)3 * nn) != 0;
    for (int d = 0; d < 3 && !bad; ++d)
        bad = cudaMemcpy(op->points.ptr + (size_t)d * nn, points[d], (size_t)nn * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess;
    if (bad) {
        delete op;
        return nullptr;
    }
    return op;
}

}  // namespace b200fem
