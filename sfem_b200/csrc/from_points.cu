// Operators that take the node coordinates instead of cached geometry factors -- the reference's non-default paths:
//
//   tet4_laplacian_apply (operators/tet4/tet4_laplacian.cpp:70-135): FFF of every element recomputed from the fp32 points
//     in scalar_t (tet4_fff_s, operators/tet4/tet4_inline_cpu.hpp:168-222), used when the Op holds no FFF cache
//     (frontend/ops/sfem_Laplacian.cpp:43-49);
//   hex8_linear_elasticity_apply (operators/hex8/hex8_linear_elasticity.cpp:86-324): the isoparametric variant selected by
//     SFEM_HEX8_ASSUME_AFFINE=0 -- trilinear geometry, Jacobian at each of the 2x2x2 Gauss points (1-D tables :55-71).
//
// Both are one thread per element with RED.ADD scatter: they are fallbacks, the cached-geometry plans are the fast path.
#include <string.h>

#include <algorithm>
#include <vector>

#include "coloring.cuh"
#include "common.cuh"
#include "hex8_le_core.cuh"
#include "hex8_local.cuh"
#include "hex8_mass_local.cuh"

namespace b200fem {

template <typename T>
__global__ void __launch_bounds__(256)
    tet4_laplacian_points_kernel(const ptrdiff_t ne, const ptrdiff_t nn, const int32_t *__restrict__ elements,
                                 const float *__restrict__ points, const T *__restrict__ u, T *__restrict__ y) {
    for (ptrdiff_t e = (ptrdiff_t)blockIdx.x * blockDim.x + threadIdx.x; e < ne; e += (ptrdiff_t)gridDim.x * blockDim.x) {
        int32_t ev[4];
#pragma unroll
        for (int v = 0; v < 4; ++v) ev[v] = ldg_stream(&elements[v * ne + e]);
        // J[r][c] = x_r(vertex c + 1) - x_r(vertex 0); adj = det J^-1; FFF = adj adj^T / (6 det)
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
        const T sc = T(1) / (T(6) * det);
        T f[6];
        int k = 0;
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = i; j < 3; ++j) f[k++] = (A[i * 3] * A[j * 3] + A[i * 3 + 1] * A[j * 3 + 1] + A[i * 3 + 2] * A[j * 3 + 2]) * sc;
        const T u0 = u[ev[0]];
        const T d1 = u[ev[1]] - u0, d2 = u[ev[2]] - u0, d3 = u[ev[3]] - u0;
        const T w0 = f[0] * d1 + f[1] * d2 + f[2] * d3;
        const T w1 = f[1] * d1 + f[3] * d2 + f[4] * d3;
        const T w2 = f[2] * d1 + f[4] * d2 + f[5] * d3;
        red_add(&y[ev[0]], -(w0 + w1 + w2));
        red_add(&y[ev[1]], w0);
        red_add(&y[ev[2]], w1);
        red_add(&y[ev[3]], w2);
    }
}

// isoparametric Q1 hex, 2x2x2 Gauss on [0,1]^3 (weights 1/8): out_i^a += sum_q w sum_c P[a][c] dN_i/dxi_c,
// P = sigma(H) adj^T, H = G adj / det, G[b][c] = sum_i u_i^b dN_i/dxi_c, sigma = lambda tr(H) I + mu (H + H^T)
template <typename T>
__global__ void __launch_bounds__(128)
    hex8_le_isoparametric_kernel(const ptrdiff_t ne, const ptrdiff_t nn, const int32_t *__restrict__ elements,
                                 const float *__restrict__ points, const T mu, const T lambda, const T *__restrict__ u,
                                 T *__restrict__ y) {
    // 1-D shape values at the two Gauss points: hex8_q2_shape_1d (hex8_linear_elasticity.cpp:55-61)
    const T s_lo = T(0.78867513459481287), s_hi = T(0.21132486540518708);
    for (ptrdiff_t e = (ptrdiff_t)blockIdx.x * blockDim.x + threadIdx.x; e < ne; e += (ptrdiff_t)gridDim.x * blockDim.x) {
        int32_t ev[8];  // lexicographic corners
#pragma unroll
        for (int v = 0; v < 8; ++v) ev[kVtkToLex[v]] = ldg_stream(&elements[v * ne + e]);
        T X[3][8], U[3][8], O[3][8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                X[d][i] = (T)points[(size_t)d * nn + ev[i]];
                U[d][i] = u[(ptrdiff_t)ev[i] * 3 + d];
                O[d][i] = T(0);
            }
        }
#pragma unroll 1
        for (int q = 0; q < 8; ++q) {
            const int qx = q & 1, qy = (q >> 1) & 1, qz = (q >> 2) & 1;
            // shape value of corner bit b at Gauss point index g: (b == g) ? s_lo : s_hi  (point 0 sits near corner 0)
            T gx[8], gy[8], gz[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int bx = i & 1, by = (i >> 1) & 1, bz = (i >> 2) & 1;
                const T sx = (bx == qx) ? s_lo : s_hi, sy = (by == qy) ? s_lo : s_hi, sz = (bz == qz) ? s_lo : s_hi;
                gx[i] = (bx ? T(1) : T(-1)) * sy * sz;
                gy[i] = sx * (by ? T(1) : T(-1)) * sz;
                gz[i] = sx * sy * (bz ? T(1) : T(-1));
            }
            T J[9], G[9];
#pragma unroll
            for (int k = 0; k < 9; ++k) J[k] = G[k] = T(0);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    J[r * 3 + 0] = le_fma(X[r][i], gx[i], J[r * 3 + 0]);
                    J[r * 3 + 1] = le_fma(X[r][i], gy[i], J[r * 3 + 1]);
                    J[r * 3 + 2] = le_fma(X[r][i], gz[i], J[r * 3 + 2]);
                    G[r * 3 + 0] = le_fma(U[r][i], gx[i], G[r * 3 + 0]);
                    G[r * 3 + 1] = le_fma(U[r][i], gy[i], G[r * 3 + 1]);
                    G[r * 3 + 2] = le_fma(U[r][i], gz[i], G[r * 3 + 2]);
                }
            }
            T A[9];  // adj[c][k] = det * d xi_c / d x_k
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
            const T idet = T(1) / det;
            T Hm[9];  // H[b][k] = sum_c G[b][c] adj[c][k] / det
#pragma unroll
            for (int b = 0; b < 3; ++b)
#pragma unroll
                for (int k = 0; k < 3; ++k) Hm[b * 3 + k] = (G[b * 3] * A[k] + G[b * 3 + 1] * A[3 + k] + G[b * 3 + 2] * A[6 + k]) * idet;
            const T ltr = lambda * (Hm[0] + Hm[4] + Hm[8]);
            T S[9];
#pragma unroll
            for (int a = 0; a < 3; ++a)
#pragma unroll
                for (int k = 0; k < 3; ++k) S[a * 3 + k] = mu * (Hm[a * 3 + k] + Hm[k * 3 + a]) + (a == k ? ltr : T(0));
            const T w = T(0.125);
            T P[9];  // P[a][c] = w sum_k S[a][k] adj[c][k]   (the determinant of dV cancels the 1 / det of the test gradient)
#pragma unroll
            for (int a = 0; a < 3; ++a)
#pragma unroll
                for (int c = 0; c < 3; ++c) P[a * 3 + c] = w * (S[a * 3] * A[c * 3] + S[a * 3 + 1] * A[c * 3 + 1] + S[a * 3 + 2] * A[c * 3 + 2]);
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int a = 0; a < 3; ++a) O[a][i] += P[a * 3] * gx[i] + P[a * 3 + 1] * gy[i] + P[a * 3 + 2] * gz[i];
        }
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int a = 0; a < 3; ++a) red_add(&y[(ptrdiff_t)ev[i] * 3 + a], O[a][i]);
    }
}

// tet4 linear elasticity from the cached adjugate / determinant: tet4_linear_elasticity_apply_opt
// (operators/tet4/tet4_linear_elasticity.cpp:237-300; local kernel tet4_linear_elasticity_apply_adj,
// tet4_linear_elasticity_inline_cpu.hpp:8-90). Constant strain: H = G adj / det with G[b][c] = u_b(v_{c+1}) - u_b(v_0),
// P = (lambda tr(H) I + mu (H + H^T)) adj^T / 6, out_b(v_{c+1}) = P[b][c], out_b(v_0) = -sum_c P[b][c].
template <typename T>
__device__ __forceinline__ void tet4_le_local(const T A[9], const T det, const T mu, const T lambda, const T (&u)[4][3], T (&o)[4][3]) {
    const T idet = T(1) / det;
    T H[9];
#pragma unroll
    for (int b = 0; b < 3; ++b) {
        const T g0 = u[1][b] - u[0][b], g1 = u[2][b] - u[0][b], g2 = u[3][b] - u[0][b];
#pragma unroll
        for (int k = 0; k < 3; ++k) H[b * 3 + k] = idet * (A[k] * g0 + A[3 + k] * g1 + A[6 + k] * g2);
    }
    const T sixth = T(1.0 / 6.0);
    const T ltr = lambda * (H[0] + H[4] + H[8]);
    T S[9];
#pragma unroll
    for (int b = 0; b < 3; ++b)
#pragma unroll
        for (int k = 0; k < 3; ++k) S[b * 3 + k] = sixth * (mu * (H[b * 3 + k] + H[k * 3 + b]) + (b == k ? ltr : T(0)));
#pragma unroll
    for (int b = 0; b < 3; ++b) {
        T sum = T(0);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const T p = A[c * 3] * S[b * 3] + A[c * 3 + 1] * S[b * 3 + 1] + A[c * 3 + 2] * S[b * 3 + 2];
            o[c + 1][b] = p;
            sum += p;
        }
        o[0][b] = -sum;
    }
}

template <typename T>
__global__ void __launch_bounds__(256)
    tet4_le_apply_kernel(const ptrdiff_t ne, const int32_t *__restrict__ elements, const float *__restrict__ adjugate,
                         const float *__restrict__ determinant, const T mu, const T lambda, const T *__restrict__ u, T *__restrict__ y) {
    for (ptrdiff_t e = (ptrdiff_t)blockIdx.x * blockDim.x + threadIdx.x; e < ne; e += (ptrdiff_t)gridDim.x * blockDim.x) {
        int32_t ev[4];
#pragma unroll
        for (int v = 0; v < 4; ++v) ev[v] = ldg_stream(&elements[v * ne + e]);
        T A[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) A[k] = (T)ldg_stream(&adjugate[k * ne + e]);
        const T det = (T)ldg_stream(&determinant[e]);
        T ul[4][3], ol[4][3];
#pragma unroll
        for (int v = 0; v < 4; ++v)
#pragma unroll
            for (int b = 0; b < 3; ++b) ul[v][b] = u[(ptrdiff_t)ev[v] * 3 + b];
        tet4_le_local(A, det, mu, lambda, ul, ol);
#pragma unroll
        for (int v = 0; v < 4; ++v)
#pragma unroll
            for (int b = 0; b < 3; ++b) red_add(&y[(ptrdiff_t)ev[v] * 3 + b], ol[v][b]);
    }
}

// diagonal: row (v, b) of the 12 x 12 element matrix applied to its own unit vector
template <typename T>
__global__ void __launch_bounds__(256)
    tet4_le_diag_kernel(const ptrdiff_t ne, const int32_t *__restrict__ elements, const float *__restrict__ adjugate,
                        const float *__restrict__ determinant, const T mu, const T lambda, T *__restrict__ diag) {
    for (ptrdiff_t e = (ptrdiff_t)blockIdx.x * blockDim.x + threadIdx.x; e < ne; e += (ptrdiff_t)gridDim.x * blockDim.x) {
        T A[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) A[k] = (T)adjugate[k * ne + e];
        const T det = (T)determinant[e];
        for (int v = 0; v < 4; ++v) {
            for (int b = 0; b < 3; ++b) {
                T ul[4][3], ol[4][3];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int c = 0; c < 3; ++c) ul[i][c] = T(0);
                ul[v][b] = T(1);
                tet4_le_local(A, det, mu, lambda, ul, ol);
                red_add(&diag[(ptrdiff_t)elements[v * ne + e] * 3 + b], ol[v][b]);
            }
        }
    }
}

// Symmetric 3 x 3 node blocks of the tet4 elasticity matrix, 6 entries (0,0) (0,1) (0,2) (1,1) (1,2) (2,2) per node, AoS [nn * 6]:
// tet4_linear_elasticity_block_diag_sym (operators/tet4/tet4_linear_elasticity.cpp:360-440). Row b of the block of corner v is the
// response at (v, :) to the unit vector (v, b). The reference calls its point kernel with weight 1 instead of the 1/6 of the
// reference volume, so ITS blocks are 6 x the blocks of its own apply / diag; `scale` = 6 reproduces that (the default of the
// plans, option "reference_block_diag"), 1 gives the consistent blocks. FROM_POINTS: geometry recomputed like the reference does.
template <typename T, bool FROM_POINTS>
__global__ void __launch_bounds__(256)
    tet4_le_block_diag_kernel(const ptrdiff_t ne, const ptrdiff_t nn, const int32_t *__restrict__ elements, const float *__restrict__ points,
                              const float *__restrict__ adjugate, const float *__restrict__ determinant, const T mu, const T lambda,
                              const T scale, T *__restrict__ out6) {
    for (ptrdiff_t e = (ptrdiff_t)blockIdx.x * blockDim.x + threadIdx.x; e < ne; e += (ptrdiff_t)gridDim.x * blockDim.x) {
        int32_t ev[4];
#pragma unroll
        for (int v = 0; v < 4; ++v) ev[v] = elements[v * ne + e];
        T A[9], det;
        if (FROM_POINTS) {
            T J[9];
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                const T p0 = (T)points[(size_t)r * nn + ev[0]];
#pragma unroll
                for (int c = 0; c < 3; ++c) J[r * 3 + c] = (T)points[(size_t)r * nn + ev[c + 1]] - p0;
            }
            A[0] = J[4] * J[8] - J[5] * J[7];
            A[1] = J[2] * J[7] - J[1] * J[8];
            A[2] = J[1] * J[5] - J[2] * J[4];
            A[3] = J[5] * J[6] - J[3] * J[8];
            A[4] = J[0] * J[8] - J[2] * J[6];
            A[5] = J[2] * J[3] - J[0] * J[5];
            A[6] = J[3] * J[7] - J[4] * J[6];
            A[7] = J[1] * J[6] - J[0] * J[7];
            A[8] = J[0] * J[4] - J[1] * J[3];
            det = J[0] * A[0] + J[1] * A[3] + J[2] * A[6];
        } else {
#pragma unroll
            for (int k = 0; k < 9; ++k) A[k] = (T)adjugate[k * ne + e];
            det = (T)determinant[e];
        }
        for (int v = 0; v < 4; ++v) {
            T blk[3][3];
#pragma unroll
            for (int b = 0; b < 3; ++b) {
                T ul[4][3], ol[4][3];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int c = 0; c < 3; ++c) ul[i][c] = (i == v && c == b) ? T(1) : T(0);
                tet4_le_local(A, det, mu, lambda, ul, ol);
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    T r = T(0);
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        if (i == v) r = ol[i][c];
                    blk[b][c] = r;
                }
            }
            T *dst = &out6[(ptrdiff_t)ev[v] * 6];
            red_add(&dst[0], scale * blk[0][0]);
            red_add(&dst[1], scale * blk[0][1]);
            red_add(&dst[2], scale * blk[0][2]);
            red_add(&dst[3], scale * blk[1][1]);
            red_add(&dst[4], scale * blk[1][2]);
            red_add(&dst[5], scale * blk[2][2]);
        }
    }
}

// tet4_linear_elasticity_apply / _diag (operators/tet4/tet4_linear_elasticity.cpp:86-160, 162-230): adjugate and determinant
// recomputed from the points in the vector precision (tet4_adjugate_and_det_s), then the same local operator
template <typename T, bool DIAG>
__global__ void __launch_bounds__(256)
    tet4_le_points_kernel(const ptrdiff_t ne, const ptrdiff_t nn, const int32_t *__restrict__ elements, const float *__restrict__ points,
                          const T mu, const T lambda, const T *__restrict__ u, T *__restrict__ y) {
    for (ptrdiff_t e = (ptrdiff_t)blockIdx.x * blockDim.x + threadIdx.x; e < ne; e += (ptrdiff_t)gridDim.x * blockDim.x) {
        int32_t ev[4];
#pragma unroll
        for (int v = 0; v < 4; ++v) ev[v] = ldg_stream(&elements[v * ne + e]);
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
        T ul[4][3], ol[4][3];
        if (DIAG) {
            for (int v = 0; v < 4; ++v) {
                for (int b = 0; b < 3; ++b) {
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int c = 0; c < 3; ++c) ul[i][c] = T(0);
                    ul[v][b] = T(1);
                    tet4_le_local(A, det, mu, lambda, ul, ol);
                    red_add(&y[(ptrdiff_t)ev[v] * 3 + b], ol[v][b]);
                }
            }
        } else {
#pragma unroll
            for (int v = 0; v < 4; ++v)
#pragma unroll
                for (int b = 0; b < 3; ++b) ul[v][b] = u[(ptrdiff_t)ev[v] * 3 + b];
            tet4_le_local(A, det, mu, lambda, ul, ol);
#pragma unroll
            for (int v = 0; v < 4; ++v)
#pragma unroll
                for (int b = 0; b < 3; ++b) red_add(&y[(ptrdiff_t)ev[v] * 3 + b], ol[v][b]);
        }
    }
}

struct Tet4LinearElasticity final : OpBase {
    DeviceArray<int32_t> elements;  // [4][ne]
    DeviceArray<float> adjugate;    // SoA [9][ne]
    DeviceArray<float> determinant;
    double mu = 1, lambda = 1;
    template <typename T>
    int apply_t(const T *x, T *y, cudaStream_t s) {
        if (ne == 0) return B200FEM_SUCCESS;
        const int grid = (int)std::min<ptrdiff_t>((ne + 255) / 256, (ptrdiff_t)sm_count() * 16);
        tet4_le_apply_kernel<T><<<grid, 256, 0, s>>>(ne, elements.ptr, adjugate.ptr, determinant.ptr, (T)mu, (T)lambda, x, y);
        B2_CHECK_LAUNCH();
        return B200FEM_SUCCESS;
    }
    int apply(const void *x, void *y, bool overwrite, cudaStream_t s) override {
        if (overwrite) B2_CUDA(cudaMemsetAsync(y, 0, (size_t)nn * 3 * real_type, s));
        return real_type == B200FEM_F64 ? apply_t<double>((const double *)x, (double *)y, s) : apply_t<float>((const float *)x, (float *)y, s);
    }
    int diag(void *d, cudaStream_t s) override {
        if (ne == 0) return B200FEM_SUCCESS;
        const int grid = (int)std::min<ptrdiff_t>((ne + 255) / 256, (ptrdiff_t)sm_count() * 16);
        if (real_type == B200FEM_F64)
            tet4_le_diag_kernel<double><<<grid, 256, 0, s>>>(ne, elements.ptr, adjugate.ptr, determinant.ptr, mu, lambda, (double *)d);
        else
            tet4_le_diag_kernel<float><<<grid, 256, 0, s>>>(ne, elements.ptr, adjugate.ptr, determinant.ptr, (float)mu, (float)lambda, (float *)d);
        B2_CHECK_LAUNCH();
        return B200FEM_SUCCESS;
    }
    int reference_block_diag = 1;  // reproduce the reference's 6 x blocks (see tet4_le_block_diag_kernel)
    int block_diag_sym(void *d, cudaStream_t s) override {
        if (ne == 0) return B200FEM_SUCCESS;
        const int grid = (int)std::min<ptrdiff_t>((ne + 255) / 256, (ptrdiff_t)sm_count() * 16);
        const double sc = reference_block_diag ? 6.0 : 1.0;
        if (real_type == B200FEM_F64)
            tet4_le_block_diag_kernel<double, false><<<grid, 256, 0, s>>>(ne, nn, elements.ptr, nullptr, adjugate.ptr, determinant.ptr, mu, lambda, sc,
                                                                        (double *)d);
        else
            tet4_le_block_diag_kernel<float, false><<<grid, 256, 0, s>>>(ne, nn, elements.ptr, nullptr, adjugate.ptr, determinant.ptr, (float)mu,
                                                                       (float)lambda, (float)sc, (float *)d);
        B2_CHECK_LAUNCH();
        return B200FEM_SUCCESS;
    }
    int set_option(const char *key, int value) override {
        if (!strcmp(key, "reference_block_diag")) { reference_block_diag = value; return B200FEM_SUCCESS; }
        return OpBase::set_option(key, value);
    }
    int get_option(const char *key, int *value) override {
        if (!strcmp(key, "reference_block_diag")) { *value = reference_block_diag; return B200FEM_SUCCESS; }
        return OpBase::get_option(key, value);
    }
    int launches(bool overwrite) const override { return 1 + (overwrite ? 1 : 0); }
};

OpBase *make_tet4_linear_elasticity(ptrdiff_t ne, ptrdiff_t nn, const int32_t *const *elements, const float *adjugate,
                                    const float *determinant, double mu, double lambda, int real_type) {
    if (!check_indices(4, ne, nn, elements)) {
        set_error("tet4 linear elasticity: node index out of range");
        return nullptr;
    }
    auto *op = new Tet4LinearElasticity();
    op->elem_type = B200FEM_TET4;
    op->real_type = real_type;
    op->block_size = 3;
    op->ne = ne;
    op->nn = nn;
    op->nxe = 4;
    op->geom_bytes_per_elem = 40;
    op->mu = mu;
    op->lambda = lambda;
    std::vector<int32_t> el((size_t)4 * ne);
    for (int v = 0; v < 4; ++v) std::copy(elements[v], elements[v] + ne, el.begin() + (size_t)v * ne);
    std::vector<float> a((size_t)9 * ne);
    for (ptrdiff_t e = 0; e < ne; ++e)
        for (int k = 0; k < 9; ++k) a[(size_t)k * ne + e] = adjugate[e * 9 + k];
    if (op->elements.upload(el.data(), el.size()) || op->adjugate.upload(a.data(), a.size()) || op->determinant.upload(determinant, (size_t)ne)) {
        delete op;
        return nullptr;
    }
    return op;
}

// Mass (consistent: y += M x) and lumped mass of the isoparametric Q1 hex: hex8_apply_mass / hex8_assemble_lumped_mass
// (operators/hex8/hex8_mass.cpp:8-103); the 27-point element operator is hex8_mass_local (hex8_mass_local.cuh)
template <typename T, bool LUMPED>
__global__ void __launch_bounds__(128)
    hex8_mass_kernel(const ptrdiff_t ne, const ptrdiff_t nn, const int32_t *__restrict__ elements, const float *__restrict__ points,
                     const T *__restrict__ u, T *__restrict__ y) {
    for (ptrdiff_t e = (ptrdiff_t)blockIdx.x * blockDim.x + threadIdx.x; e < ne; e += (ptrdiff_t)gridDim.x * blockDim.x) {
        int32_t ev[8];  // lexicographic corners
#pragma unroll
        for (int v = 0; v < 8; ++v) ev[kVtkToLex[v]] = ldg_stream(&elements[v * ne + e]);
        T X[3][8], U[8], O[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
#pragma unroll
            for (int d = 0; d < 3; ++d) X[d][i] = (T)points[(size_t)d * nn + ev[i]];
            U[i] = LUMPED ? T(1) : u[ev[i]];
        }
        hex8_mass_local(X, U, O);
#pragma unroll
        for (int i = 0; i < 8; ++i) red_add(&y[ev[i]], O[i]);
    }
}

// lumped mass of the P1 tet: det / 24 on each of its nodes (tet4_assemble_lumped_mass, operators/tet4/tet4_mass.cpp:66-92,
// 231-282; the reference expands the determinant into its 24 monomials, here it is the triple product of the edge vectors).
// NXE = 10: the row sums of the P2 mass matrix on the affine tet of the 4 corners, -det / 120 on a corner and det / 30 on an
// edge node (tet10_assemble_lumped_mass -> tet10_hrz_lumped_mass_kernel, operators/tet10/tet10_mass.cpp:443-583: the active
// code of that routine is the plain row sum, negative on the corners; the HRZ-scaled variant above it is commented out)
template <typename T, int NXE>
__global__ void __launch_bounds__(256)
    tet_lumped_mass_kernel(const ptrdiff_t ne, const ptrdiff_t nn, const int32_t *__restrict__ elements, const float *__restrict__ points,
                           T *__restrict__ y) {
    for (ptrdiff_t e = (ptrdiff_t)blockIdx.x * blockDim.x + threadIdx.x; e < ne; e += (ptrdiff_t)gridDim.x * blockDim.x) {
        int32_t ev[NXE];
#pragma unroll
        for (int v = 0; v < NXE; ++v) ev[v] = ldg_stream(&elements[v * ne + e]);
        T p[3][4];
#pragma unroll
        for (int d = 0; d < 3; ++d)
#pragma unroll
            for (int v = 0; v < 4; ++v) p[d][v] = (T)points[(size_t)d * nn + ev[v]];
        const T ax = p[0][1] - p[0][0], ay = p[1][1] - p[1][0], az = p[2][1] - p[2][0];
        const T bx = p[0][2] - p[0][0], by = p[1][2] - p[1][0], bz = p[2][2] - p[2][0];
        const T cx = p[0][3] - p[0][0], cy = p[1][3] - p[1][0], cz = p[2][3] - p[2][0];
        const T det = ax * (by * cz - bz * cy) - ay * (bx * cz - bz * cx) + az * (bx * cy - by * cx);
        if (NXE == 4) {
            const T m = det * T(1.0 / 24.0);
#pragma unroll
            for (int v = 0; v < 4; ++v) red_add(&y[ev[v]], m);
        } else {
            const T mc = -det * T(1.0 / 120.0), me = det * T(1.0 / 30.0);
#pragma unroll
            for (int v = 0; v < NXE; ++v) red_add(&y[ev[v]], v < 4 ? mc : me);
        }
    }
}

struct PointsOp final : OpBase {
    DeviceArray<int32_t> elements;  // SoA [nxe][ne]
    DeviceArray<float> points;      // [3][nn]
    double mu = 1, lambda = 1;
    bool elasticity = false;
    bool mass = false;  // Mass / LumpedMass: apply = hex8_apply_mass (tet4: none, like the reference's dispatcher), diag = lumped mass

    template <typename T>
    int apply_t(const T *x, T *y, cudaStream_t s) {
        if (ne == 0) return B200FEM_SUCCESS;
        if (mass) {
            if (nxe != 8) B2_FAIL("mass apply: hex8 only (apply_mass, operators/mass.cpp:158-207, has no tet4 / tet10 case)");
            const int grid = (int)std::min<ptrdiff_t>((ne + 127) / 128, (ptrdiff_t)sm_count() * 16);
            hex8_mass_kernel<T, false><<<grid, 128, 0, s>>>(ne, nn, elements.ptr, points.ptr, x, y);
        } else if (elasticity && nxe == 4) {
            const int grid = (int)std::min<ptrdiff_t>((ne + 255) / 256, (ptrdiff_t)sm_count() * 16);
            tet4_le_points_kernel<T, false><<<grid, 256, 0, s>>>(ne, nn, elements.ptr, points.ptr, (T)mu, (T)lambda, x, y);
        } else if (elasticity) {
            const int grid = (int)std::min<ptrdiff_t>((ne + 127) / 128, (ptrdiff_t)sm_count() * 16);
            hex8_le_isoparametric_kernel<T><<<grid, 128, 0, s>>>(ne, nn, elements.ptr, points.ptr, (T)mu, (T)lambda, x, y);
        } else {
            const int grid = (int)std::min<ptrdiff_t>((ne + 255) / 256, (ptrdiff_t)sm_count() * 16);
            tet4_laplacian_points_kernel<T><<<grid, 256, 0, s>>>(ne, nn, elements.ptr, points.ptr, x, y);
        }
        B2_CHECK_LAUNCH();
        return B200FEM_SUCCESS;
    }
    int apply(const void *x, void *y, bool overwrite, cudaStream_t s) override {
        if (overwrite) B2_CUDA(cudaMemsetAsync(y, 0, (size_t)nn * block_size * real_type, s));
        return real_type == B200FEM_F64 ? apply_t<double>((const double *)x, (double *)y, s) : apply_t<float>((const float *)x, (float *)y, s);
    }
    int diag(void *d, cudaStream_t s) override {
        if (mass) {  // LumpedMass::hessian_diag: diag += row sums of the mass matrix
            if (ne == 0) return B200FEM_SUCCESS;
            const bool f64 = real_type == B200FEM_F64;
            if (nxe == 8) {
                const int grid = (int)std::min<ptrdiff_t>((ne + 127) / 128, (ptrdiff_t)sm_count() * 16);
                if (f64)
                    hex8_mass_kernel<double, true><<<grid, 128, 0, s>>>(ne, nn, elements.ptr, points.ptr, nullptr, (double *)d);
                else
                    hex8_mass_kernel<float, true><<<grid, 128, 0, s>>>(ne, nn, elements.ptr, points.ptr, nullptr, (float *)d);
            } else {
                const int grid = (int)std::min<ptrdiff_t>((ne + 255) / 256, (ptrdiff_t)sm_count() * 16);
                if (nxe == 10 && f64)
                    tet_lumped_mass_kernel<double, 10><<<grid, 256, 0, s>>>(ne, nn, elements.ptr, points.ptr, (double *)d);
                else if (nxe == 10)
                    tet_lumped_mass_kernel<float, 10><<<grid, 256, 0, s>>>(ne, nn, elements.ptr, points.ptr, (float *)d);
                else if (f64)
                    tet_lumped_mass_kernel<double, 4><<<grid, 256, 0, s>>>(ne, nn, elements.ptr, points.ptr, (double *)d);
                else
                    tet_lumped_mass_kernel<float, 4><<<grid, 256, 0, s>>>(ne, nn, elements.ptr, points.ptr, (float *)d);
            }
            B2_CHECK_LAUNCH();
            return B200FEM_SUCCESS;
        }
        if (!(elasticity && nxe == 4)) return OpBase::diag(d, s);
        if (ne == 0) return B200FEM_SUCCESS;
        const int grid = (int)std::min<ptrdiff_t>((ne + 255) / 256, (ptrdiff_t)sm_count() * 16);
        if (real_type == B200FEM_F64)
            tet4_le_points_kernel<double, true><<<grid, 256, 0, s>>>(ne, nn, elements.ptr, points.ptr, mu, lambda, nullptr, (double *)d);
        else
            tet4_le_points_kernel<float, true><<<grid, 256, 0, s>>>(ne, nn, elements.ptr, points.ptr, (float)mu, (float)lambda, nullptr, (float *)d);
        B2_CHECK_LAUNCH();
        return B200FEM_SUCCESS;
    }
    int reference_block_diag = 1;
    int block_diag_sym(void *d, cudaStream_t s) override {
        if (!(elasticity && nxe == 4)) return OpBase::block_diag_sym(d, s);
        if (ne == 0) return B200FEM_SUCCESS;
        const int grid = (int)std::min<ptrdiff_t>((ne + 255) / 256, (ptrdiff_t)sm_count() * 16);
        const double sc = reference_block_diag ? 6.0 : 1.0;
        if (real_type == B200FEM_F64)
            tet4_le_block_diag_kernel<double, true><<<grid, 256, 0, s>>>(ne, nn, elements.ptr, points.ptr, nullptr, nullptr, mu, lambda, sc, (double *)d);
        else
            tet4_le_block_diag_kernel<float, true><<<grid, 256, 0, s>>>(ne, nn, elements.ptr, points.ptr, nullptr, nullptr, (float)mu, (float)lambda,
                                                                      (float)sc, (float *)d);
        B2_CHECK_LAUNCH();
        return B200FEM_SUCCESS;
    }
    int set_option(const char *key, int value) override {
        if (!strcmp(key, "reference_block_diag")) { reference_block_diag = value; return B200FEM_SUCCESS; }
        return OpBase::set_option(key, value);
    }
    int get_option(const char *key, int *value) override {
        if (!strcmp(key, "reference_block_diag")) { *value = reference_block_diag; return B200FEM_SUCCESS; }
        return OpBase::get_option(key, value);
    }
    int launches(bool overwrite) const override { return 1 + (overwrite ? 1 : 0); }
};

static OpBase *make_points_op(int nxe, ptrdiff_t ne, ptrdiff_t nn, const int32_t *const *elements, const float *const *points,
                              int real_type, const char *what) {
    if (!check_indices(nxe, ne, nn, elements)) {
        set_error(std::string(what) + ": node index out of range");
        return nullptr;
    }
    auto *op = new PointsOp();
    op->real_type = real_type;
    op->ne = ne;
    op->nn = nn;
    op->nxe = nxe;
    op->geom_bytes_per_elem = nxe * 3 * 4;  // the corner coordinates an element reads
    std::vector<int32_t> el((size_t)nxe * ne);
    for (int v = 0; v < nxe; ++v) std::copy(elements[v], elements[v] + ne, el.begin() + (size_t)v * ne);
    bool bad = op->elements.upload(el.data(), el.size()) != 0 || op->points.alloc((size_t)3 * nn) != 0;
    for (int d = 0; d < 3 && !bad; ++d)
        bad = cudaMemcpy(op->points.ptr + (size_t)d * nn, points[d], (size_t)nn * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess;
    if (bad) {
        delete op;
        return nullptr;
    }
    return op;
}

OpBase *make_tet4_laplacian_from_points(ptrdiff_t ne, ptrdiff_t nn, const int32_t *const *elements, const float *const *points,
                                        int real_type) {
    OpBase *op = make_points_op(4, ne, nn, elements, points, real_type, "tet4 laplacian (from points)");
    if (op) op->elem_type = B200FEM_TET4;
    return op;
}

OpBase *make_mass_from_points(int element_type, ptrdiff_t ne, ptrdiff_t nn, const int32_t *const *elements, const float *const *points,
                              int real_type) {
    OpBase *op = make_points_op(element_type == B200FEM_HEX8 ? 8 : element_type == B200FEM_TET10 ? 10 : 4, ne, nn, elements, points, real_type, "mass");
    if (op) {
        op->elem_type = element_type;
        static_cast<PointsOp *>(op)->mass = true;
    }
    return op;
}

OpBase *make_tet4_le_from_points(ptrdiff_t ne, ptrdiff_t nn, const int32_t *const *elements, const float *const *points, double mu,
                                 double lambda, int real_type) {
    OpBase *op = make_points_op(4, ne, nn, elements, points, real_type, "tet4 linear elasticity (from points)");
    if (op) {
        auto *p = static_cast<PointsOp *>(op);
        p->elem_type = B200FEM_TET4;
        p->block_size = 3;
        p->elasticity = true;
        p->mu = mu;
        p->lambda = lambda;
    }
    return op;
}

OpBase *make_hex8_le_isoparametric(ptrdiff_t ne, ptrdiff_t nn, const int32_t *const *elements, const float *const *points, double mu,
                                   double lambda, int real_type) {
    OpBase *op = make_points_op(8, ne, nn, elements, points, real_type, "hex8 linear elasticity (isoparametric)");
    if (op) {
        auto *p = static_cast<PointsOp *>(op);
        p->elem_type = B200FEM_HEX8;
        p->block_size = 3;
        p->elasticity = true;
        p->mu = mu;
        p->lambda = lambda;
    }
    return op;
}

}  // namespace b200fem
