/* oracle_port.c -- plain-C CPU restatement of the reference's matrix-free apply path.
 *
 * TEST INFRASTRUCTURE ONLY. Nothing under sfem_b200/ links, loads or calls this file; only tests/,
 * __graft_entry__.smoke() and the cpu_baseline / --impl reference legs of bench.py may (and they prefer
 * oracle/_ref, the reference's own sources compiled out of tree, when it is present).
 *
 * Every function has the name and signature of the reference function it restates, so oracle/ref.py binds
 * either library. The restatement is textbook finite elements (Gauss quadrature of the weak forms), NOT a
 * transcription of the reference's generated closed forms; tests/test_oracle.py pins it against oracle/_ref on
 * random inputs and against the committed vectors in tests/golden/.
 *
 *   tet4_fff_fill                      operators/tet4/tet4_fff.cpp:4-27 (FFF = J^-1 J^-T det/6, fp32 arithmetic)
 *   tet4_laplacian_apply_opt / _diag_opt   operators/tet4/tet4_laplacian.cpp:237-322
 *   hex8_fff_fill                      operators/hex8/hex8_fff.cpp:5-37      (Jacobian at (1/2,1/2,1/2))
 *   hex8_adjugate_and_det_fill         operators/hex8/hex8_jacobian.cpp:5-38
 *   hex8_laplacian_apply_opt           operators/hex8/hex8_laplacian.cpp:309-345
 *   affine_hex8_linear_elasticity_apply_adjugate_soa   operators/hex8/hex8_linear_elasticity.cpp:625-787
 *   affine_sshex8_laplacian_stencil_apply_fff          operators/sshex8/sshex8_laplacian.cpp:698-769
 *   affine_sshex8_linear_elasticity_apply_macro_adjugate  operators/sshex8/sshex8_linear_elasticity.cpp:275-424
 *
 * Types follow cmake/SFEMNumbers.cmake:5-14: idx_t int32, geom_t / jacobian_t fp32, real_t fp64
 * (-DORACLE_REAL_FLOAT32: fp32).
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef int32_t idx_t;
typedef float geom_t;
typedef float jacobian_t;
#ifdef ORACLE_REAL_FLOAT32
typedef float real_t;
#else
typedef double real_t;
#endif
typedef real_t scalar_t;

int oracle_ref_sizeof_real(void) { return (int)sizeof(real_t); }

/* ---------------------------------------------------------------- small dense helpers */
static void adjugate3(const double J[9], double A[9], double *det) {
    A[0] = J[4] * J[8] - J[5] * J[7];
    A[1] = J[2] * J[7] - J[1] * J[8];
    A[2] = J[1] * J[5] - J[2] * J[4];
    A[3] = J[5] * J[6] - J[3] * J[8];
    A[4] = J[0] * J[8] - J[2] * J[6];
    A[5] = J[2] * J[3] - J[0] * J[5];
    A[6] = J[3] * J[7] - J[4] * J[6];
    A[7] = J[1] * J[6] - J[0] * J[7];
    A[8] = J[0] * J[4] - J[1] * J[3];
    *det = J[0] * A[0] + J[1] * A[3] + J[2] * A[6];
}

/* VTK hex8 corner signs */
static const int HX[8] = {-1, 1, 1, -1, -1, 1, 1, -1};
static const int HY[8] = {-1, -1, 1, 1, -1, -1, 1, 1};
static const int HZ[8] = {-1, -1, -1, -1, 1, 1, 1, 1};

/* reference-space gradient of the 8 trilinear shape functions on [0,1]^3 at (x,y,z) */
static void hex8_ref_grad(double x, double y, double z, double g[8][3]) {
    for (int v = 0; v < 8; ++v) {
        const double fx = HX[v] > 0 ? x : 1 - x, fy = HY[v] > 0 ? y : 1 - y, fz = HZ[v] > 0 ? z : 1 - z;
        g[v][0] = HX[v] * fy * fz;
        g[v][1] = fx * HY[v] * fz;
        g[v][2] = fx * fy * HZ[v];
    }
}

static const double GQ[2] = {0.5 - 0.28867513459481288225, 0.5 + 0.28867513459481288225}; /* 2-pt Gauss on [0,1] */

/* 8x8 Laplacian matrix for a constant symmetric FFF (exactly integrated by 2x2x2 Gauss) */
static void hex8_laplacian_matrix(const scalar_t fff[6], scalar_t K[8][8]) {
    const scalar_t F[3][3] = {{fff[0], fff[1], fff[2]}, {fff[1], fff[3], fff[4]}, {fff[2], fff[4], fff[5]}};
    memset(K, 0, sizeof(scalar_t) * 64);
    for (int a = 0; a < 2; ++a)
        for (int b = 0; b < 2; ++b)
            for (int c = 0; c < 2; ++c) {
                double g[8][3];
                hex8_ref_grad(GQ[a], GQ[b], GQ[c], g);
                for (int i = 0; i < 8; ++i) {
                    scalar_t Fg[3];
                    for (int r = 0; r < 3; ++r) Fg[r] = F[r][0] * (scalar_t)g[i][0] + F[r][1] * (scalar_t)g[i][1] + F[r][2] * (scalar_t)g[i][2];
                    for (int j = 0; j < 8; ++j)
                        K[i][j] += (scalar_t)0.125 * (Fg[0] * (scalar_t)g[j][0] + Fg[1] * (scalar_t)g[j][1] + Fg[2] * (scalar_t)g[j][2]);
                }
            }
}

/* 24x24 elasticity matrix, dof order [ux0..7 | uy0..7 | uz0..7], from adjugate (row j = det * grad xi_j) and det */
static void hex8_elasticity_matrix(const scalar_t adj[9], const scalar_t det, const scalar_t mu, const scalar_t lambda,
                                   scalar_t K[24][24]) {
    memset(K, 0, sizeof(scalar_t) * 24 * 24);
    for (int a = 0; a < 2; ++a)
        for (int b = 0; b < 2; ++b)
            for (int c = 0; c < 2; ++c) {
                double g[8][3];
                hex8_ref_grad(GQ[a], GQ[b], GQ[c], g);
                scalar_t G[8][3]; /* physical gradients: G[v][k] = sum_j g[v][j] adj[j][k] / det */
                for (int v = 0; v < 8; ++v)
                    for (int k = 0; k < 3; ++k)
                        G[v][k] = ((scalar_t)g[v][0] * adj[0 * 3 + k] + (scalar_t)g[v][1] * adj[1 * 3 + k] + (scalar_t)g[v][2] * adj[2 * 3 + k]) / det;
                const scalar_t w = (scalar_t)0.125 * det;
                for (int i = 0; i < 8; ++i)
                    for (int j = 0; j < 8; ++j) {
                        const scalar_t gg = G[i][0] * G[j][0] + G[i][1] * G[j][1] + G[i][2] * G[j][2];
                        for (int p = 0; p < 3; ++p)
                            for (int q = 0; q < 3; ++q) {
                                scalar_t v = lambda * G[i][p] * G[j][q] + mu * G[i][q] * G[j][p];
                                if (p == q) v += mu * gg;
                                K[p * 8 + i][q * 8 + j] += w * v;
                            }
                    }
            }
}

/* ---------------------------------------------------------------- geometry */
int tet4_fff_fill(const ptrdiff_t ne, idx_t **const elements, geom_t **const points, jacobian_t *const fff) {
#pragma omp parallel for
    for (ptrdiff_t e = 0; e < ne; ++e) {
        /* the reference evaluates this in geom_t (fp32) */
        float J[9];
        const idx_t i0 = elements[0][e];
        for (int c = 0; c < 3; ++c) {
            const idx_t ic = elements[c + 1][e];
            for (int r = 0; r < 3; ++r) J[r * 3 + c] = points[r][ic] - points[r][i0];
        }
        float A[9];
        A[0] = J[4] * J[8] - J[5] * J[7];
        A[1] = J[2] * J[7] - J[1] * J[8];
        A[2] = J[1] * J[5] - J[2] * J[4];
        A[3] = J[5] * J[6] - J[3] * J[8];
        A[4] = J[0] * J[8] - J[2] * J[6];
        A[5] = J[2] * J[3] - J[0] * J[5];
        A[6] = J[3] * J[7] - J[4] * J[6];
        A[7] = J[1] * J[6] - J[0] * J[7];
        A[8] = J[0] * J[4] - J[1] * J[3];
        const float det = J[0] * A[0] + J[1] * A[3] + J[2] * A[6];
        int k = 0;
        for (int i = 0; i < 3; ++i)
            for (int j = i; j < 3; ++j)
                fff[e * 6 + k++] = (A[i * 3] * A[j * 3] + A[i * 3 + 1] * A[j * 3 + 1] + A[i * 3 + 2] * A[j * 3 + 2]) / (6 * det);
    }
    return 0;
}

static void hex8_center_jacobian(const ptrdiff_t e, idx_t **const elements, geom_t **const points, double J[9]) {
    for (int k = 0; k < 9; ++k) J[k] = 0;
    for (int v = 0; v < 8; ++v) {
        const idx_t i = elements[v][e];
        for (int r = 0; r < 3; ++r) {
            J[r * 3 + 0] += 0.25 * HX[v] * (double)points[r][i];
            J[r * 3 + 1] += 0.25 * HY[v] * (double)points[r][i];
            J[r * 3 + 2] += 0.25 * HZ[v] * (double)points[r][i];
        }
    }
}

int hex8_fff_fill(const ptrdiff_t ne, idx_t **const elements, geom_t **const points, jacobian_t *const fff) {
#pragma omp parallel for
    for (ptrdiff_t e = 0; e < ne; ++e) {
        double J[9], A[9], det;
        hex8_center_jacobian(e, elements, points, J);
        adjugate3(J, A, &det);
        int k = 0;
        for (int i = 0; i < 3; ++i)
            for (int j = i; j < 3; ++j)
                fff[e * 6 + k++] = (jacobian_t)((A[i * 3] * A[j * 3] + A[i * 3 + 1] * A[j * 3 + 1] + A[i * 3 + 2] * A[j * 3 + 2]) / det);
    }
    return 0;
}

int hex8_adjugate_and_det_fill(const ptrdiff_t ne, idx_t **const elements, geom_t **const points, jacobian_t *const adjugate,
                               jacobian_t *const determinant) {
#pragma omp parallel for
    for (ptrdiff_t e = 0; e < ne; ++e) {
        double J[9], A[9], det;
        hex8_center_jacobian(e, elements, points, J);
        adjugate3(J, A, &det);
        for (int k = 0; k < 9; ++k) adjugate[e * 9 + k] = (jacobian_t)A[k];
        determinant[e] = (jacobian_t)det;
    }
    return 0;
}

/* ---------------------------------------------------------------- tet4 Laplacian */
static void tet4_matrix(const scalar_t f[6], scalar_t K[4][4]) {
    const scalar_t F[3][3] = {{f[0], f[1], f[2]}, {f[1], f[3], f[4]}, {f[2], f[4], f[5]}};
    const scalar_t g[4][3] = {{-1, -1, -1}, {1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            scalar_t s = 0;
            for (int r = 0; r < 3; ++r)
                for (int c = 0; c < 3; ++c) s += g[i][r] * F[r][c] * g[j][c];
            K[i][j] = s;
        }
}

int tet4_laplacian_apply_opt(const ptrdiff_t ne, idx_t **const elements, const jacobian_t *const fff, const real_t *const u,
                             real_t *const values) {
#pragma omp parallel for
    for (ptrdiff_t e = 0; e < ne; ++e) {
        scalar_t f[6], K[4][4];
        for (int k = 0; k < 6; ++k) f[k] = (scalar_t)fff[e * 6 + k];
        tet4_matrix(f, K);
        idx_t ev[4];
        for (int v = 0; v < 4; ++v) ev[v] = elements[v][e];
        for (int i = 0; i < 4; ++i) {
            scalar_t s = 0;
            for (int j = 0; j < 4; ++j) s += K[i][j] * u[ev[j]];
#pragma omp atomic update
            values[ev[i]] += s;
        }
    }
    return 0;
}

int tet4_laplacian_diag_opt(const ptrdiff_t ne, idx_t **const elements, const jacobian_t *const fff, real_t *const diag) {
#pragma omp parallel for
    for (ptrdiff_t e = 0; e < ne; ++e) {
        scalar_t f[6], K[4][4];
        for (int k = 0; k < 6; ++k) f[k] = (scalar_t)fff[e * 6 + k];
        tet4_matrix(f, K);
        for (int i = 0; i < 4; ++i) {
#pragma omp atomic update
            diag[elements[i][e]] += K[i][i];
        }
    }
    return 0;
}

/* ---------------------------------------------------------------- hex8 */
int hex8_laplacian_apply_opt(const ptrdiff_t ne, idx_t **const elements, const jacobian_t *const fff, const real_t *const u,
                             real_t *const values) {
#pragma omp parallel for
    for (ptrdiff_t e = 0; e < ne; ++e) {
        scalar_t f[6], K[8][8];
        for (int k = 0; k < 6; ++k) f[k] = (scalar_t)fff[e * 6 + k];
        hex8_laplacian_matrix(f, K);
        idx_t ev[8];
        for (int v = 0; v < 8; ++v) ev[v] = elements[v][e];
        for (int i = 0; i < 8; ++i) {
            scalar_t s = 0;
            for (int j = 0; j < 8; ++j) s += K[i][j] * u[ev[j]];
#pragma omp atomic update
            values[ev[i]] += s;
        }
    }
    return 0;
}

int affine_hex8_linear_elasticity_apply_adjugate_soa(const ptrdiff_t ne, const ptrdiff_t nnodes, idx_t **const elements,
                                                     jacobian_t **const adjugate /* [9][ne] */, const jacobian_t *const determinant,
                                                     const real_t mu, const real_t lambda, const ptrdiff_t u_stride,
                                                     const real_t *const ux, const real_t *const uy, const real_t *const uz,
                                                     const ptrdiff_t out_stride, real_t *const outx, real_t *const outy,
                                                     real_t *const outz) {
    (void)nnodes;
    const real_t *uu[3] = {ux, uy, uz};
    real_t *oo[3] = {outx, outy, outz};
#pragma omp parallel for
    for (ptrdiff_t e = 0; e < ne; ++e) {
        scalar_t adj[9];
        for (int k = 0; k < 9; ++k) adj[k] = (scalar_t)adjugate[k][e];
        static _Thread_local scalar_t K[24][24];
        hex8_elasticity_matrix(adj, (scalar_t)determinant[e], mu, lambda, K);
        scalar_t ue[24], oe[24];
        for (int p = 0; p < 3; ++p)
            for (int v = 0; v < 8; ++v) ue[p * 8 + v] = uu[p][(ptrdiff_t)elements[v][e] * u_stride];
        for (int i = 0; i < 24; ++i) {
            scalar_t s = 0;
            for (int j = 0; j < 24; ++j) s += K[i][j] * ue[j];
            oe[i] = s;
        }
        for (int p = 0; p < 3; ++p)
            for (int v = 0; v < 8; ++v) {
#pragma omp atomic update
                oo[p][(ptrdiff_t)elements[v][e] * out_stride] += oe[p * 8 + v];
            }
    }
    return 0;
}

/* ---------------------------------------------------------------- sshex8 (macro element = (L+1)^3 lattice) */
static inline int lidx(int L, int x, int y, int z) { return z * (L + 1) * (L + 1) + y * (L + 1) + x; }

static void micro_corners(int L, int x, int y, int z, int lev[8]) {
    lev[0] = lidx(L, x, y, z);
    lev[1] = lidx(L, x + 1, y, z);
    lev[2] = lidx(L, x + 1, y + 1, z);
    lev[3] = lidx(L, x, y + 1, z);
    lev[4] = lidx(L, x, y, z + 1);
    lev[5] = lidx(L, x + 1, y, z + 1);
    lev[6] = lidx(L, x + 1, y + 1, z + 1);
    lev[7] = lidx(L, x, y + 1, z + 1);
}

int affine_sshex8_laplacian_stencil_apply_fff(const int level, const ptrdiff_t ne, idx_t **const elements,
                                              const jacobian_t *const fff, const real_t *const u, real_t *const values) {
    const int L = level, nxe = (L + 1) * (L + 1) * (L + 1);
#pragma omp parallel
    {
        scalar_t *eu = (scalar_t *)malloc(sizeof(scalar_t) * nxe), *ev = (scalar_t *)malloc(sizeof(scalar_t) * nxe);
#pragma omp for
        for (ptrdiff_t e = 0; e < ne; ++e) {
            scalar_t f[6], K[8][8];
            for (int k = 0; k < 6; ++k) f[k] = (scalar_t)fff[e * 6 + k] * ((scalar_t)1 / (scalar_t)L); /* micro FFF = FFF h */
            hex8_laplacian_matrix(f, K);
            for (int n = 0; n < nxe; ++n) {
                eu[n] = u[elements[n][e]];
                ev[n] = 0;
            }
            for (int z = 0; z < L; ++z)
                for (int y = 0; y < L; ++y)
                    for (int x = 0; x < L; ++x) {
                        int lev[8];
                        micro_corners(L, x, y, z, lev);
                        for (int i = 0; i < 8; ++i) {
                            scalar_t s = 0;
                            for (int j = 0; j < 8; ++j) s += K[i][j] * eu[lev[j]];
                            ev[lev[i]] += s;
                        }
                    }
            for (int n = 0; n < nxe; ++n) {
#pragma omp atomic update
                values[elements[n][e]] += ev[n];
            }
        }
        free(eu);
        free(ev);
    }
    return 0;
}

int affine_sshex8_linear_elasticity_apply_macro_adjugate(const int level, const ptrdiff_t ne, const ptrdiff_t nnodes,
                                                         idx_t **const elements, geom_t **const points /* unused */,
                                                         const jacobian_t *const macro_adjugate /* [ne*9] */,
                                                         const jacobian_t *const macro_determinant, const real_t mu,
                                                         const real_t lambda, const ptrdiff_t u_stride, const real_t *const ux,
                                                         const real_t *const uy, const real_t *const uz, const ptrdiff_t out_stride,
                                                         real_t *const outx, real_t *const outy, real_t *const outz) {
    (void)nnodes;
    (void)points;
    const int L = level, nxe = (L + 1) * (L + 1) * (L + 1);
    const real_t *uu[3] = {ux, uy, uz};
    real_t *oo[3] = {outx, outy, outz};
    const scalar_t h = (scalar_t)1 / (scalar_t)L;
#pragma omp parallel
    {
        scalar_t *eu = (scalar_t *)malloc(sizeof(scalar_t) * nxe * 3), *ev = (scalar_t *)malloc(sizeof(scalar_t) * nxe * 3);
        scalar_t(*K)[24] = (scalar_t(*)[24])malloc(sizeof(scalar_t) * 24 * 24);
#pragma omp for
        for (ptrdiff_t e = 0; e < ne; ++e) {
            scalar_t adj[9];
            for (int k = 0; k < 9; ++k) adj[k] = (scalar_t)macro_adjugate[e * 9 + k] * h * h; /* adj(J h) = adj(J) h^2 */
            hex8_elasticity_matrix(adj, (scalar_t)macro_determinant[e] * h * h * h, mu, lambda, K);
            for (int p = 0; p < 3; ++p)
                for (int n = 0; n < nxe; ++n) {
                    eu[p * nxe + n] = uu[p][(ptrdiff_t)elements[n][e] * u_stride];
                    ev[p * nxe + n] = 0;
                }
            for (int z = 0; z < L; ++z)
                for (int y = 0; y < L; ++y)
                    for (int x = 0; x < L; ++x) {
                        int lev[8];
                        micro_corners(L, x, y, z, lev);
                        scalar_t ul[24];
                        for (int p = 0; p < 3; ++p)
                            for (int v = 0; v < 8; ++v) ul[p * 8 + v] = eu[p * nxe + lev[v]];
                        for (int i = 0; i < 24; ++i) {
                            scalar_t s = 0;
                            for (int j = 0; j < 24; ++j) s += K[i][j] * ul[j];
                            ev[(i / 8) * nxe + lev[i % 8]] += s;
                        }
                    }
            for (int p = 0; p < 3; ++p)
                for (int n = 0; n < nxe; ++n) {
#pragma omp atomic update
                    oo[p][(ptrdiff_t)elements[n][e] * out_stride] += ev[p * nxe + n];
                }
        }
        free(eu);
        free(ev);
        free(K);
    }
    return 0;
}
