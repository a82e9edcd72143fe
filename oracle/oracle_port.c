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
 *   tet10_laplacian_apply_opt / _diag_opt, macro_tet4_laplacian_apply_opt / _diag_opt
 *                                      operators/tet10/tet10_laplacian.cpp:232-302, operators/macro_tet4/macro_tet4_laplacian.cpp:401-490
 *   tet10_linear_elasticity_apply_opt, macro_tet4_linear_elasticity_apply_opt
 *                                      operators/tet10/tet10_linear_elasticity.cpp:209-283, operators/macro_tet4/macro_tet4_linear_elasticity.cpp:206-277
 *   packed_laplacian_apply_u16         the pack loop of PackedLaplacianApply::apply / ::apply_two_pass,
 *                                      frontend/ops/sfem_PackedLaplacian.cpp:111-262 (needs smesh: not compilable into oracle/_ref)
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

/* ---------------------------------------------------------------- 10-node tetrahedra (P2 and P1-iso-P2 Laplacian)
 * tet10_laplacian_apply_opt / _diag_opt     operators/tet10/tet10_laplacian.cpp:232-302
 * macro_tet4_laplacian_apply_opt            operators/macro_tet4/macro_tet4_laplacian.cpp:401-440
 * Textbook restatement: the 10 x 10 P2 stiffness matrix by the 4-point degree-2 Gauss rule of the tetrahedron (the integrand --
 * a product of two linear gradients -- is quadratic, so the rule is exact; the reference uses an 8-point rule that is exact too),
 * the macro element as 8 linear sub-tets on the mid-edge nodes with K_sub = G^T FFF G |T_sub| / |T| and G from the sub-tet's
 * barycentric gradients in the parent's reference coordinates. Node order: 0-3 corners, 4 = (0,1), 5 = (1,2), 6 = (0,2),
 * 7 = (0,3), 8 = (1,3), 9 = (2,3). The reference's tet10 diagonal is 1/6 of the matrix diagonal. */
static const int P2_EDGE[6][2] = {{0, 1}, {1, 2}, {0, 2}, {0, 3}, {1, 3}, {2, 3}};
static const double P2_GL[4][3] = {{-1, -1, -1}, {1, 0, 0}, {0, 1, 0}, {0, 0, 1}}; /* grad lambda_a */

static void tet10_matrix(const scalar_t f[6], scalar_t K[10][10]) {
    const double F[3][3] = {{f[0], f[1], f[2]}, {f[1], f[3], f[4]}, {f[2], f[4], f[5]}};
    const double a = 0.58541019662496845446, b = 0.13819660112501051518;
    memset(K, 0, sizeof(scalar_t) * 100);
    double Kd[10][10] = {{0}};
    for (int q = 0; q < 4; ++q) {
        double lam[4] = {b, b, b, b};
        lam[q] = a;
        double g[10][3];
        for (int v = 0; v < 4; ++v)
            for (int d = 0; d < 3; ++d) g[v][d] = (4 * lam[v] - 1) * P2_GL[v][d];
        for (int k = 0; k < 6; ++k)
            for (int d = 0; d < 3; ++d)
                g[4 + k][d] = 4 * (lam[P2_EDGE[k][0]] * P2_GL[P2_EDGE[k][1]][d] + lam[P2_EDGE[k][1]] * P2_GL[P2_EDGE[k][0]][d]);
        for (int i = 0; i < 10; ++i)
            for (int j = 0; j < 10; ++j) {
                double s = 0;
                for (int r = 0; r < 3; ++r)
                    for (int c = 0; c < 3; ++c) s += g[i][r] * F[r][c] * g[j][c];
                Kd[i][j] += 0.25 * s; /* weights sum to 1: FFF already carries det / 6 */
            }
    }
    for (int i = 0; i < 10; ++i)
        for (int j = 0; j < 10; ++j) K[i][j] = (scalar_t)Kd[i][j];
}

static void macro_tet4_matrix(const scalar_t f[6], scalar_t K[10][10]) {
    static const int sub[8][4] = {{0, 4, 6, 7}, {4, 1, 5, 8}, {6, 5, 2, 9}, {7, 8, 9, 3}, {4, 5, 6, 8}, {7, 4, 6, 8}, {6, 5, 9, 8}, {7, 6, 9, 8}};
    static const double P[10][3] = {{0, 0, 0},    {1, 0, 0},   {0, 1, 0},   {0, 0, 1},   {.5, 0, 0},
                                    {.5, .5, 0},  {0, .5, 0},  {0, 0, .5},  {.5, 0, .5}, {0, .5, .5}};
    const double F[3][3] = {{f[0], f[1], f[2]}, {f[1], f[3], f[4]}, {f[2], f[4], f[5]}};
    double Kd[10][10] = {{0}};
    for (int s = 0; s < 8; ++s) {
        /* A = [p1 - p0 | p2 - p0 | p3 - p0]; gradients of the sub-tet's barycentric coordinates = rows of A^-1 */
        double A[9], Ai[9], det;
        for (int c = 0; c < 3; ++c)
            for (int r = 0; r < 3; ++r) A[r * 3 + c] = P[sub[s][c + 1]][r] - P[sub[s][0]][r];
        adjugate3(A, Ai, &det);
        double g[4][3];
        for (int c = 0; c < 3; ++c) {
            for (int d = 0; d < 3; ++d) g[c + 1][d] = Ai[c * 3 + d] / det;
        }
        for (int d = 0; d < 3; ++d) g[0][d] = -(g[1][d] + g[2][d] + g[3][d]);
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 4; ++j) {
                double v = 0;
                for (int r = 0; r < 3; ++r)
                    for (int c = 0; c < 3; ++c) v += g[i][r] * F[r][c] * g[j][c];
                Kd[sub[s][i]][sub[s][j]] += v * fabs(det); /* |T_sub| / |T_parent| = |det A| (both reference volumes are 1/6) */
            }
    }
    for (int i = 0; i < 10; ++i)
        for (int j = 0; j < 10; ++j) K[i][j] = (scalar_t)Kd[i][j];
}

static int p2_laplacian_apply(const int macro, const ptrdiff_t ne, idx_t **const elements, const jacobian_t *const fff,
                              const real_t *const u, real_t *const values) {
#pragma omp parallel for
    for (ptrdiff_t e = 0; e < ne; ++e) {
        scalar_t f[6], K[10][10];
        for (int k = 0; k < 6; ++k) f[k] = (scalar_t)fff[e * 6 + k];
        if (macro)
            macro_tet4_matrix(f, K);
        else
            tet10_matrix(f, K);
        idx_t ev[10];
        for (int v = 0; v < 10; ++v) ev[v] = elements[v][e];
        for (int i = 0; i < 10; ++i) {
            scalar_t s = 0;
            for (int j = 0; j < 10; ++j) s += K[i][j] * u[ev[j]];
#pragma omp atomic update
            values[ev[i]] += s;
        }
    }
    return 0;
}
int tet10_laplacian_apply_opt(const ptrdiff_t ne, idx_t **const elements, const jacobian_t *const fff, const real_t *const u,
                              real_t *const values) {
    return p2_laplacian_apply(0, ne, elements, fff, u, values);
}
int macro_tet4_laplacian_apply_opt(const ptrdiff_t ne, idx_t **const elements, const jacobian_t *const fff, const real_t *const u,
                                   real_t *const values) {
    return p2_laplacian_apply(1, ne, elements, fff, u, values);
}
static int p2_laplacian_diag(const int macro, const ptrdiff_t ne, idx_t **const elements, const jacobian_t *const fff, real_t *const diag) {
#pragma omp parallel for
    for (ptrdiff_t e = 0; e < ne; ++e) {
        scalar_t f[6], K[10][10];
        for (int k = 0; k < 6; ++k) f[k] = (scalar_t)fff[e * 6 + k];
        if (macro)
            macro_tet4_matrix(f, K);
        else
            tet10_matrix(f, K);
        for (int i = 0; i < 10; ++i) {
            const scalar_t d = macro ? K[i][i] : K[i][i] / 6; /* the reference's tet10 diagonal: 1/6 of the matrix diagonal */
#pragma omp atomic update
            diag[elements[i][e]] += d;
        }
    }
    return 0;
}
int tet10_laplacian_diag_opt(const ptrdiff_t ne, idx_t **const elements, const jacobian_t *const fff, real_t *const diag) {
    return p2_laplacian_diag(0, ne, elements, fff, diag);
}
int macro_tet4_laplacian_diag_opt(const ptrdiff_t ne, idx_t **const elements, const jacobian_t *const fff, real_t *const diag) {
    return p2_laplacian_diag(1, ne, elements, fff, diag);
}

/* ---------------------------------------------------------------- PackedLaplacian
 * packed_laplacian_apply<uint16_t> -> PackedLaplacianApply<pack_idx_t, NXE, MicroKernel>::apply and ::apply_two_pass
 * (frontend/ops/sfem_PackedLaplacian.cpp:111-176, 178-262, dispatch 551-700). The reference file needs smesh (absent) and cannot be
 * compiled into oracle/_ref, so the pack loop is restated here as it is written there: per pack, copy the owned node range and
 * gather the ghosts into a pack-local vector, run the element operator on the 16-bit local connectivity into a pack-local result,
 * add the result back (owned range, then ghosts). Two-pass: owned range added in place, ghost results parked in a buffer and
 * reduced per destination node through the CSR lists. element_type: 4 tet4, 8 hex8, 10 tet10, 100 macro_tet4 (smesh::ElemType). */
static void packed_element(const int element_type, const scalar_t f[6], const scalar_t *in, const uint16_t *const *el, const ptrdiff_t e,
                           scalar_t *out) {
    if (element_type == 4) {
        scalar_t K[4][4];
        tet4_matrix(f, K);
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 4; ++j) out[el[i][e]] += K[i][j] * in[el[j][e]];
    } else if (element_type == 8) {
        scalar_t K[8][8];
        hex8_laplacian_matrix(f, K);
        for (int i = 0; i < 8; ++i)
            for (int j = 0; j < 8; ++j) out[el[i][e]] += K[i][j] * in[el[j][e]];
    } else {
        scalar_t K[10][10];
        if (element_type == 10)
            tet10_matrix(f, K);
        else
            macro_tet4_matrix(f, K);
        for (int i = 0; i < 10; ++i)
            for (int j = 0; j < 10; ++j) out[el[i][e]] += K[i][j] * in[el[j][e]];
    }
}

int packed_laplacian_apply_u16(const int element_type, const ptrdiff_t n_packs, const ptrdiff_t n_elements_per_pack, const ptrdiff_t n_elements,
                               const ptrdiff_t max_nodes_per_pack, const uint16_t *const *elements, const jacobian_t *const fff,
                               const ptrdiff_t *owned_nodes_ptr, const ptrdiff_t *n_shared_nodes, const ptrdiff_t *ghost_ptr,
                               const idx_t *ghost_idx, const real_t *const u, real_t *const values,
                               /* two-pass lists, all NULL for the one-pass form */
                               const ptrdiff_t n_ghost_reduce_rows, const ptrdiff_t *ghost_reduce_ptr, const ptrdiff_t *ghost_reduce_idx,
                               const idx_t *ghost_reduce_dest) {
    (void)n_shared_nodes; /* sequential: the owned tail needs no atomics here */
    if (element_type != 4 && element_type != 8 && element_type != 10 && element_type != 100) return 1;
    const int two_pass = ghost_reduce_ptr != NULL;
    scalar_t *in = (scalar_t *)malloc(sizeof(scalar_t) * (size_t)(max_nodes_per_pack > 0 ? max_nodes_per_pack : 1));
    scalar_t *out = (scalar_t *)calloc((size_t)(max_nodes_per_pack > 0 ? max_nodes_per_pack : 1), sizeof(scalar_t));
    real_t *ghost_buf = two_pass ? (real_t *)calloc((size_t)(ghost_ptr[n_packs] > 0 ? ghost_ptr[n_packs] : 1), sizeof(real_t)) : NULL;
    for (ptrdiff_t p = 0; p < n_packs; ++p) {
        const ptrdiff_t e0 = p * n_elements_per_pack, e1 = (p + 1) * n_elements_per_pack < n_elements ? (p + 1) * n_elements_per_pack : n_elements;
        const ptrdiff_t n_contiguous = owned_nodes_ptr[p + 1] - owned_nodes_ptr[p], n_ghost = ghost_ptr[p + 1] - ghost_ptr[p];
        const idx_t *ghosts = ghost_idx + ghost_ptr[p];
        for (ptrdiff_t k = 0; k < n_contiguous; ++k) in[k] = u[owned_nodes_ptr[p] + k];
        for (ptrdiff_t k = 0; k < n_ghost; ++k) in[n_contiguous + k] = u[ghosts[k]];
        for (ptrdiff_t e = e0; e < e1; ++e) {
            scalar_t f[6];
            for (int k = 0; k < 6; ++k) f[k] = (scalar_t)fff[e * 6 + k];
            packed_element(element_type, f, in, elements, e, out);
        }
        for (ptrdiff_t k = 0; k < n_contiguous; ++k) {
            values[owned_nodes_ptr[p] + k] += out[k];
            out[k] = 0;
        }
        for (ptrdiff_t k = 0; k < n_ghost; ++k) {
            if (two_pass)
                ghost_buf[ghost_ptr[p] + k] = out[n_contiguous + k];
            else
                values[ghosts[k]] += out[n_contiguous + k];
            out[n_contiguous + k] = 0;
        }
    }
    if (two_pass)
        for (ptrdiff_t r = 0; r < n_ghost_reduce_rows; ++r) {
            real_t sum = 0;
            for (ptrdiff_t j = ghost_reduce_ptr[r]; j < ghost_reduce_ptr[r + 1]; ++j) sum += ghost_buf[ghost_reduce_idx[j]];
            values[ghost_reduce_dest[r]] += sum;
        }
    free(in);
    free(out);
    free(ghost_buf);
    return 0;
}

/* ---------------------------------------------------------------- 10-node tetrahedra, linear elasticity
 * tet10_linear_elasticity_apply_opt        operators/tet10/tet10_linear_elasticity.cpp:209-283
 * macro_tet4_linear_elasticity_apply_opt   operators/macro_tet4/macro_tet4_linear_elasticity.cpp:206-277
 * Textbook restatement: physical shape gradients grad phi_i = J^-T grad_ref phi_i = adjugate^T grad_ref phi_i / det, bilinear form
 * a(u, v) = int lambda div(u) div(v) + 2 mu eps(u) : eps(v); tet10 by the exact 4-point degree-2 rule (the reference's 8-point rule
 * carries face centroids written as 0.333333333333 and is 1.5e-12 off), macro_tet4 as 8 constant-strain sub-tets. u / out AoS-strided
 * component pointers like the reference's signature. */
static void p2_elasticity_element(const int macro, const scalar_t adj[9], const scalar_t det, const scalar_t mu, const scalar_t lambda,
                                  const double u[10][3], double out[10][3]) {
    static const int sub[8][4] = {{0, 4, 6, 7}, {4, 1, 5, 8}, {6, 5, 2, 9}, {7, 8, 9, 3}, {4, 5, 6, 8}, {7, 4, 6, 8}, {6, 5, 9, 8}, {7, 6, 9, 8}};
    static const double P[10][3] = {{0, 0, 0},    {1, 0, 0},   {0, 1, 0},   {0, 0, 1},   {.5, 0, 0},
                                    {.5, .5, 0},  {0, .5, 0},  {0, 0, .5},  {.5, 0, .5}, {0, .5, .5}};
    const double a = 0.58541019662496845446, b = 0.13819660112501051518;
    for (int i = 0; i < 10; ++i)
        for (int c = 0; c < 3; ++c) out[i][c] = 0;
    const int npieces = macro ? 8 : 4;
    for (int q = 0; q < npieces; ++q) {
        double gref[10][3] = {{0}}; /* reference gradients of the shape functions active on this piece */
        double w;                   /* piece weight relative to the parent volume det / 6 */
        if (!macro) {
            double lam[4] = {b, b, b, b};
            lam[q] = a;
            for (int v = 0; v < 4; ++v)
                for (int d = 0; d < 3; ++d) gref[v][d] = (4 * lam[v] - 1) * P2_GL[v][d];
            for (int k = 0; k < 6; ++k)
                for (int d = 0; d < 3; ++d)
                    gref[4 + k][d] = 4 * (lam[P2_EDGE[k][0]] * P2_GL[P2_EDGE[k][1]][d] + lam[P2_EDGE[k][1]] * P2_GL[P2_EDGE[k][0]][d]);
            w = 0.25;
        } else {
            double A[9], Ai[9], dA;
            for (int c = 0; c < 3; ++c)
                for (int r = 0; r < 3; ++r) A[r * 3 + c] = P[sub[q][c + 1]][r] - P[sub[q][0]][r];
            adjugate3(A, Ai, &dA);
            for (int c = 0; c < 3; ++c)
                for (int d = 0; d < 3; ++d) gref[sub[q][c + 1]][d] = Ai[c * 3 + d] / dA;
            for (int d = 0; d < 3; ++d) gref[sub[q][0]][d] = -(gref[sub[q][1]][d] + gref[sub[q][2]][d] + gref[sub[q][3]][d]);
            w = fabs(dA);
        }
        /* physical gradients: g_phys[d] = sum_j adj[j][d] gref[j] / det   (adjugate = det J^-1, row j = det grad xi_j) */
        double g[10][3];
        for (int i = 0; i < 10; ++i)
            for (int d = 0; d < 3; ++d) g[i][d] = (adj[0 * 3 + d] * gref[i][0] + adj[1 * 3 + d] * gref[i][1] + adj[2 * 3 + d] * gref[i][2]) / det;
        double H[3][3] = {{0}}; /* displacement gradient */
        for (int i = 0; i < 10; ++i)
            for (int c = 0; c < 3; ++c)
                for (int d = 0; d < 3; ++d) H[c][d] += u[i][c] * g[i][d];
        const double tr = H[0][0] + H[1][1] + H[2][2];
        double S[3][3];
        for (int c = 0; c < 3; ++c)
            for (int d = 0; d < 3; ++d) S[c][d] = mu * (H[c][d] + H[d][c]) + (c == d ? lambda * tr : 0);
        const double vol = w * det / 6.0;
        for (int i = 0; i < 10; ++i)
            for (int c = 0; c < 3; ++c) out[i][c] += vol * (S[c][0] * g[i][0] + S[c][1] * g[i][1] + S[c][2] * g[i][2]);
    }
}

static int p2_elasticity_apply(const int macro, const ptrdiff_t ne, idx_t **const elements, const jacobian_t *const adjugate,
                               const jacobian_t *const determinant, const real_t mu, const real_t lambda, const ptrdiff_t u_stride,
                               const real_t *const ux, const real_t *const uy, const real_t *const uz, const ptrdiff_t out_stride,
                               real_t *const outx, real_t *const outy, real_t *const outz) {
    const real_t *const uc[3] = {ux, uy, uz};
    real_t *const oc[3] = {outx, outy, outz};
    for (ptrdiff_t e = 0; e < ne; ++e) {
        scalar_t A[9];
        for (int k = 0; k < 9; ++k) A[k] = (scalar_t)adjugate[e * 9 + k];
        double ul[10][3], ol[10][3];
        for (int v = 0; v < 10; ++v)
            for (int c = 0; c < 3; ++c) ul[v][c] = uc[c][(ptrdiff_t)elements[v][e] * u_stride];
        p2_elasticity_element(macro, A, (scalar_t)determinant[e], mu, lambda, ul, ol);
        for (int v = 0; v < 10; ++v)
            for (int c = 0; c < 3; ++c) oc[c][(ptrdiff_t)elements[v][e] * out_stride] += (real_t)ol[v][c];
    }
    return 0;
}
int tet10_linear_elasticity_apply_opt(const ptrdiff_t ne, idx_t **const elements, const jacobian_t *const adjugate,
                                      const jacobian_t *const determinant, const real_t mu, const real_t lambda, const ptrdiff_t u_stride,
                                      const real_t *const ux, const real_t *const uy, const real_t *const uz, const ptrdiff_t out_stride,
                                      real_t *const outx, real_t *const outy, real_t *const outz) {
    return p2_elasticity_apply(0, ne, elements, adjugate, determinant, mu, lambda, u_stride, ux, uy, uz, out_stride, outx, outy, outz);
}
int macro_tet4_linear_elasticity_apply_opt(const ptrdiff_t ne, idx_t **const elements, const jacobian_t *const adjugate,
                                           const jacobian_t *const determinant, const real_t mu, const real_t lambda,
                                           const ptrdiff_t u_stride, const real_t *const ux, const real_t *const uy, const real_t *const uz,
                                           const ptrdiff_t out_stride, real_t *const outx, real_t *const outy, real_t *const outz) {
    return p2_elasticity_apply(1, ne, elements, adjugate, determinant, mu, lambda, u_stride, ux, uy, uz, out_stride, outx, outy, outz);
}
