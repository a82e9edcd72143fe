// CPU harness for sfem_b200/csrc/p2_tet_local.cuh: the exact per-element arithmetic of the tet10 / macro_tet4 CUDA
// kernels, run sequentially on the host so the formulation can be checked against the reference without a GPU.
#include <stddef.h>
#include <stdint.h>

#include "p2_tet_local.cuh"

using namespace b200fem;

template <typename T>
static void laplacian(int macro, int diag, int reference_scale, ptrdiff_t ne, const int32_t *const *el, const float *fff, const T *u, T *out) {
    for (ptrdiff_t e = 0; e < ne; ++e) {
        T f[6], ul[10], ol[10];
        for (int k = 0; k < 6; ++k) f[k] = (T)fff[e * 6 + k];
        for (int v = 0; v < 10; ++v) ul[v] = diag ? T(0) : u[el[v][e]];
        if (diag) {
            if (macro)
                macro_tet4_laplacian_local_diag(f, ol);
            else
                tet10_laplacian_local_diag(f, reference_scale != 0, ol);
        } else {
            if (macro)
                macro_tet4_laplacian_local(f, ul, ol);
            else
                tet10_laplacian_local(f, ul, ol);
        }
        for (int v = 0; v < 10; ++v) out[el[v][e]] += ol[v];
    }
}

template <typename T>
static void elasticity(int macro, int diag, ptrdiff_t ne, const int32_t *const *el, const float *adj, const float *det, double mu, double lambda,
                       const T *u, T *out) {
    for (ptrdiff_t e = 0; e < ne; ++e) {
        T A[9], ul[10][3], ol[10][3];
        for (int k = 0; k < 9; ++k) A[k] = (T)adj[e * 9 + k];
        const T d = (T)det[e];
        if (diag) {
            for (int v = 0; v < 10; ++v)
                for (int b = 0; b < 3; ++b) {
                    for (int i = 0; i < 10; ++i)
                        for (int c = 0; c < 3; ++c) ul[i][c] = T(0);
                    ul[v][b] = T(1);
                    if (macro)
                        p2tet_le_local<T, true>(A, d, (T)mu, (T)lambda, ul, ol);
                    else
                        p2tet_le_local<T, false>(A, d, (T)mu, (T)lambda, ul, ol);
                    out[(ptrdiff_t)el[v][e] * 3 + b] += ol[v][b];
                }
            continue;
        }
        for (int v = 0; v < 10; ++v)
            for (int b = 0; b < 3; ++b) ul[v][b] = u[(ptrdiff_t)el[v][e] * 3 + b];
        if (macro)
            p2tet_le_local<T, true>(A, d, (T)mu, (T)lambda, ul, ol);
        else
            p2tet_le_local<T, false>(A, d, (T)mu, (T)lambda, ul, ol);
        for (int v = 0; v < 10; ++v)
            for (int b = 0; b < 3; ++b) out[(ptrdiff_t)el[v][e] * 3 + b] += ol[v][b];
    }
}

extern "C" {
int p2_tet_laplacian_f64(int macro, int diag, int reference_scale, ptrdiff_t ne, const int32_t *const *el, const float *fff, const double *u, double *out) {
    laplacian<double>(macro, diag, reference_scale, ne, el, fff, u, out);
    return 0;
}
int p2_tet_laplacian_f32(int macro, int diag, int reference_scale, ptrdiff_t ne, const int32_t *const *el, const float *fff, const float *u, float *out) {
    laplacian<float>(macro, diag, reference_scale, ne, el, fff, u, out);
    return 0;
}
int p2_tet_elasticity_f64(int macro, int diag, ptrdiff_t ne, const int32_t *const *el, const float *adj, const float *det, double mu, double lambda,
                          const double *u, double *out) {
    elasticity<double>(macro, diag, ne, el, adj, det, mu, lambda, u, out);
    return 0;
}
int p2_tet_elasticity_f32(int macro, int diag, ptrdiff_t ne, const int32_t *const *el, const float *adj, const float *det, double mu, double lambda,
                          const float *u, float *out) {
    elasticity<float>(macro, diag, ne, el, adj, det, mu, lambda, u, out);
    return 0;
}
// the assembly loop of tet4_le_block_diag_kernel (sfem_b200/csrc/from_points.cu) on the host: rows of the 3 x 3 block of corner v
// from the responses to the unit vectors (v, b), 6 entries per node, scaled
int tet4_le_block_diag_f64(ptrdiff_t ne, const int32_t *const *el, const float *adj, const float *det, double mu, double lambda, double scale,
                           double *out6) {
    for (ptrdiff_t e = 0; e < ne; ++e) {
        double A[9];
        for (int k = 0; k < 9; ++k) A[k] = (double)adj[e * 9 + k];
        for (int v = 0; v < 4; ++v) {
            double blk[3][3];
            for (int b = 0; b < 3; ++b) {
                double ul[10][3] = {}, ol[10][3] = {};
                ul[v][b] = 1.0;
                tet4_le_local_add(A, (double)det[e], mu, lambda, ul, ol, 0, 1, 2, 3);
                for (int c = 0; c < 3; ++c) blk[b][c] = ol[v][c];
            }
            double *dst = out6 + (ptrdiff_t)el[v][e] * 6;
            dst[0] += scale * blk[0][0];
            dst[1] += scale * blk[0][1];
            dst[2] += scale * blk[0][2];
            dst[3] += scale * blk[1][1];
            dst[4] += scale * blk[1][2];
            dst[5] += scale * blk[2][2];
        }
    }
    return 0;
}
// the sub-tet table, for the test that re-derives it from the node positions
int p2_tet_macro_table(int s, int *nodes4, int *halfB9) {
    for (int v = 0; v < 4; ++v) nodes4[v] = MacroTet4::node(s, v);
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) halfB9[r * 3 + c] = MacroTet4::halfB(s, r, c);
    return 0;
}
}
