// Local operators of the 10-node tetrahedra: tet10 (quadratic, P2) and macro_tet4 (P1-iso-P2: 8 linear sub-tets),
// Laplacian and linear elasticity, geometry = the affine map of the 4 corners (FFF / adjugate of the parent tet4).
//
// Replaces
//   tet10_laplacian_apply_add_fff / tet10_laplacian_diag_fff   operators/tet10/tet10_laplacian_inline_cpu.hpp:94-108, 688-711
//   tet10_linear_elasticity_apply_adj / _diag_adj              operators/tet10/tet10_linear_elasticity_inline_cpu.hpp
//   macro_tet4 element_apply + tet4_sub_fff_*                  operators/macro_tet4/macro_tet4_laplacian.cpp:12-125,
//                                                              operators/macro_tet4/macro_tet4_inline_cpu.hpp:8-80
//   macro_tet4_local_apply_adj + tet4_sub_adj_*                operators/macro_tet4/macro_tet4_linear_elasticity.cpp:44-135,
//                                                              macro_tet4_inline_cpu.hpp:84-160
//
// Node order (both families): 0-3 corners, 4 = edge(0,1), 5 = (1,2), 6 = (0,2), 7 = (0,3), 8 = (1,3), 9 = (2,3).
//
// tet10. The reference integrates with an 8-point rule (4 corners, weight 1/40; 4 face centroids, weight 9/40) and
// generated per-point expressions. The gradient of a P2 function is LINEAR on the element, so everything the rule sees is
// a product of two linear functions, which it integrates exactly; we therefore never visit a quadrature point:
//     grad u (xi) = sum_k lambda_k(xi) g^(k),        g^(k) = gradient at corner k = sum_a c^(k)_a grad lambda_a,
//     c^(k)_k = 3 u_k,   c^(k)_a = 4 u_{edge(a,k)} - u_a  (a != k),
//     int lambda_k lambda_l / |T| = (1 + delta_kl) / 20,
// so with the flux at the corners w^(k) = FFF g^(k) and W^(k) = (sum_l w^(l) + w^(k)) / 20,
//     out_i = sum_k grad phi_i^(k) . W^(k),   grad phi_a^(k) = (4 delta_ak - 1) grad lambda_a,
//                                             grad phi_{ab}^(k) = 4 (delta_ak grad lambda_b + delta_bk grad lambda_a).
// 4 corner gradients + 4 fluxes instead of 8 evaluations of 10 shape-function gradients; identical in exact arithmetic.
// Linear elasticity is the same with a 3 x 3 stress per corner.
//
// macro_tet4. Sub-tet s has the reference corners p_0..p_3 (corners / edge midpoints of the parent); with
// B_s = [p_1 - p_0, p_2 - p_0, p_3 - p_0]^-1 (entries 0, +-2) its operator is the tet4 one with
//     FFF_s = B_s FFF B_s^T / 8,      adjugate_s = B_s adjugate,  determinant_s = 8 determinant
// (the reference's generated tet4_sub_fff_k / tet4_sub_adj_k are these products written out; its scaling of the
// adjugate / determinant pair is kept, the operator only sees adjugate x adjugate / determinant). The table below is
// derived from the node positions; tests/test_p2_tet_core.py re-derives it.
//
// __host__ __device__: tests/test_p2_tet_core.py runs exactly this code on the CPU against the reference.
#pragma once

#ifdef __CUDACC__
#define B2P2_HD __host__ __device__ __forceinline__
#else
#define B2P2_HD inline
#endif

namespace b200fem {

// local node on the edge between corners a and b
B2P2_HD constexpr int p2_edge_node(const int a, const int b) {
    const int lo = a < b ? a : b, hi = a < b ? b : a;
    return (lo == hi) ? lo : (lo == 0) ? (hi == 1 ? 4 : hi == 2 ? 6 : 7) : (lo == 1) ? (hi == 2 ? 5 : 8) : 9;
}

// ---------------------------------------------------------------- tet10 Laplacian
template <typename T>
B2P2_HD void tet10_laplacian_local(const T (&f)[6], const T (&u)[10], T (&out)[10]) {
    T w[4][3];
    T S[3] = {T(0), T(0), T(0)};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        T c[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) c[a] = (a == k) ? T(3) * u[k] : T(4) * u[p2_edge_node(a, k)] - u[a];
        const T g0 = c[1] - c[0], g1 = c[2] - c[0], g2 = c[3] - c[0];
        w[k][0] = f[0] * g0 + f[1] * g1 + f[2] * g2;
        w[k][1] = f[1] * g0 + f[3] * g1 + f[4] * g2;
        w[k][2] = f[2] * g0 + f[4] * g1 + f[5] * g2;
#pragma unroll
        for (int d = 0; d < 3; ++d) S[d] += w[k][d];
    }
    // d[k][a] = grad lambda_a . W^(k)
    T d[4][4], sumd[4] = {T(0), T(0), T(0), T(0)};
    const T s = T(1) / T(20);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        d[k][1] = s * (S[0] + w[k][0]);
        d[k][2] = s * (S[1] + w[k][1]);
        d[k][3] = s * (S[2] + w[k][2]);
        d[k][0] = -(d[k][1] + d[k][2] + d[k][3]);
#pragma unroll
        for (int a = 0; a < 4; ++a) sumd[a] += d[k][a];
    }
#pragma unroll
    for (int a = 0; a < 4; ++a) out[a] = T(4) * d[a][a] - sumd[a];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = a + 1; b < 4; ++b) out[p2_edge_node(a, b)] = T(4) * (d[a][b] + d[b][a]);
}

// q_ab = grad lambda_a . FFF grad lambda_b
template <typename T>
B2P2_HD void p2_fff_forms(const T (&f)[6], T (&q)[4][4]) {
    const T F[3][3] = {{f[0], f[1], f[2]}, {f[1], f[3], f[4]}, {f[2], f[4], f[5]}};
#pragma unroll
    for (int a = 1; a < 4; ++a)
#pragma unroll
        for (int b = 1; b < 4; ++b) q[a][b] = F[a - 1][b - 1];
#pragma unroll
    for (int a = 1; a < 4; ++a) q[0][a] = q[a][0] = -(F[a - 1][0] + F[a - 1][1] + F[a - 1][2]);
    q[0][0] = -(q[0][1] + q[0][2] + q[0][3]);
}

// Diagonal of the tet10 element matrix:  corner a: 3/5 q_aa;  edge (a,b): 8/5 (q_aa + q_bb + q_ab).
// The reference's tet10_laplacian_diag_fff returns ONE SIXTH of that (1/10 q_aa, 4/15 (...): generated for an FFF without
// the 1/6 of the reference volume, while tet4_fff and the apply carry it) -- reference_scale reproduces it.
template <typename T>
B2P2_HD void tet10_laplacian_local_diag(const T (&f)[6], const bool reference_scale, T (&out)[10]) {
    T q[4][4];
    p2_fff_forms(f, q);
    const T sc = reference_scale ? T(1) / T(6) : T(1);
    const T cv = sc * T(3) / T(5), ce = sc * T(8) / T(5);
#pragma unroll
    for (int a = 0; a < 4; ++a) out[a] = cv * q[a][a];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = a + 1; b < 4; ++b) out[p2_edge_node(a, b)] = ce * (q[a][a] + q[b][b] + q[a][b]);
}

// ---------------------------------------------------------------- macro_tet4 tables
struct MacroTet4 {
    // sub-tets (4 corner tets, 4 tets of the inner octahedron)
    static B2P2_HD constexpr int node(const int s, const int v) {
        constexpr int t[8][4] = {{0, 4, 6, 7}, {4, 1, 5, 8}, {6, 5, 2, 9}, {7, 8, 9, 3},
                                 {4, 5, 6, 8}, {7, 4, 6, 8}, {6, 5, 9, 8}, {7, 6, 9, 8}};
        return t[s][v];
    }
    // HALF of B_s = [p_1 - p_0 | p_2 - p_0 | p_3 - p_0]^-1, row-major (B_s itself has entries 0, +-2)
    static B2P2_HD constexpr int halfB(const int s, const int r, const int c) {
        constexpr int b[8][9] = {{1, 0, 0, 0, 1, 0, 0, 0, 1},   {1, 0, 0, 0, 1, 0, 0, 0, 1},  {1, 0, 0, 0, 1, 0, 0, 0, 1},
                                 {1, 0, 0, 0, 1, 0, 0, 0, 1},   {1, 1, 0, -1, 0, 0, 0, 0, 1}, {0, -1, -1, 0, 1, 0, 1, 1, 1},
                                 {1, 1, 0, 0, 1, 1, 0, -1, 0},  {0, 0, -1, 0, 1, 1, 1, 0, 0}};
        return b[s][r * 3 + c];
    }
};

// FFF_s = B_s FFF B_s^T / 8 = (B_s/2) FFF (B_s/2)^T / 2
template <typename T, int S>
B2P2_HD void macro_tet4_sub_fff(const T (&f)[6], T (&sf)[6]) {
    const T F[3][3] = {{f[0], f[1], f[2]}, {f[1], f[3], f[4]}, {f[2], f[4], f[5]}};
    T BF[3][3];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            T acc = T(0);
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const int b = MacroTet4::halfB(S, r, k);
                if (b == 1) acc += F[k][c];
                if (b == -1) acc -= F[k][c];
            }
            BF[r][c] = acc;
        }
    int n = 0;
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = r; c < 3; ++c) {
            T acc = T(0);
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const int b = MacroTet4::halfB(S, c, k);
                if (b == 1) acc += BF[r][k];
                if (b == -1) acc -= BF[r][k];
            }
            sf[n++] = T(0.5) * acc;
        }
}

template <typename T>
B2P2_HD void tet4_laplacian_local_add(const T (&f)[6], const T u0, const T u1, const T u2, const T u3, T &o0, T &o1, T &o2, T &o3) {
    const T d1 = u1 - u0, d2 = u2 - u0, d3 = u3 - u0;
    const T w0 = f[0] * d1 + f[1] * d2 + f[2] * d3;
    const T w1 = f[1] * d1 + f[3] * d2 + f[4] * d3;
    const T w2 = f[2] * d1 + f[4] * d2 + f[5] * d3;
    o0 -= w0 + w1 + w2;
    o1 += w0;
    o2 += w1;
    o3 += w2;
}

template <typename T, int S>
B2P2_HD void macro_tet4_laplacian_sub(const T (&f)[6], const T (&u)[10], T (&out)[10]) {
    T sf[6];
    macro_tet4_sub_fff<T, S>(f, sf);
    constexpr int n0 = MacroTet4::node(S, 0), n1 = MacroTet4::node(S, 1), n2 = MacroTet4::node(S, 2), n3 = MacroTet4::node(S, 3);
    tet4_laplacian_local_add(sf, u[n0], u[n1], u[n2], u[n3], out[n0], out[n1], out[n2], out[n3]);
}

template <typename T>
B2P2_HD void macro_tet4_laplacian_local(const T (&f)[6], const T (&u)[10], T (&out)[10]) {
#pragma unroll
    for (int i = 0; i < 10; ++i) out[i] = T(0);
    macro_tet4_laplacian_sub<T, 0>(f, u, out);
    macro_tet4_laplacian_sub<T, 1>(f, u, out);
    macro_tet4_laplacian_sub<T, 2>(f, u, out);
    macro_tet4_laplacian_sub<T, 3>(f, u, out);
    macro_tet4_laplacian_sub<T, 4>(f, u, out);
    macro_tet4_laplacian_sub<T, 5>(f, u, out);
    macro_tet4_laplacian_sub<T, 6>(f, u, out);
    macro_tet4_laplacian_sub<T, 7>(f, u, out);
}

template <typename T, int S>
B2P2_HD void macro_tet4_laplacian_diag_sub(const T (&f)[6], T (&out)[10]) {
    T sf[6];
    macro_tet4_sub_fff<T, S>(f, sf);
    out[MacroTet4::node(S, 0)] += sf[0] + T(2) * sf[1] + T(2) * sf[2] + sf[3] + T(2) * sf[4] + sf[5];
    out[MacroTet4::node(S, 1)] += sf[0];
    out[MacroTet4::node(S, 2)] += sf[3];
    out[MacroTet4::node(S, 3)] += sf[5];
}

template <typename T>
B2P2_HD void macro_tet4_laplacian_local_diag(const T (&f)[6], T (&out)[10]) {
#pragma unroll
    for (int i = 0; i < 10; ++i) out[i] = T(0);
    macro_tet4_laplacian_diag_sub<T, 0>(f, out);
    macro_tet4_laplacian_diag_sub<T, 1>(f, out);
    macro_tet4_laplacian_diag_sub<T, 2>(f, out);
    macro_tet4_laplacian_diag_sub<T, 3>(f, out);
    macro_tet4_laplacian_diag_sub<T, 4>(f, out);
    macro_tet4_laplacian_diag_sub<T, 5>(f, out);
    macro_tet4_laplacian_diag_sub<T, 6>(f, out);
    macro_tet4_laplacian_diag_sub<T, 7>(f, out);
}

// ---------------------------------------------------------------- linear elasticity
// Conventions of the tet4 operator (from_points.cu: tet4_le_local): A = adjugate (row-major, = det J^-1), reference gradient
// G[b][j] = d u_b / d xi_j, H = G A / det, S = lambda tr(H) I + mu (H + H^T), P = S A^T, out_b(i) = sum_c (grad_ref phi_i)_c P[b][c] x volume weight.

// stress-like S from a reference gradient (no volume factor)
template <typename T>
B2P2_HD void p2_le_stress(const T (&A)[9], const T idet, const T mu, const T lambda, const T (&G)[3][3], T (&S)[3][3]) {
    T H[3][3];
#pragma unroll
    for (int b = 0; b < 3; ++b)
#pragma unroll
        for (int k = 0; k < 3; ++k) H[b][k] = idet * (G[b][0] * A[k] + G[b][1] * A[3 + k] + G[b][2] * A[6 + k]);
    const T ltr = lambda * (H[0][0] + H[1][1] + H[2][2]);
#pragma unroll
    for (int b = 0; b < 3; ++b)
#pragma unroll
        for (int k = 0; k < 3; ++k) S[b][k] = mu * (H[b][k] + H[k][b]) + (b == k ? ltr : T(0));
}

// tet10: u, out as [node][component]
template <typename T>
B2P2_HD void tet10_le_local(const T (&A)[9], const T det, const T mu, const T lambda, const T (&u)[10][3], T (&out)[10][3]) {
    const T idet = T(1) / det;
    T Sk[4][3][3], Ss[3][3];
#pragma unroll
    for (int b = 0; b < 3; ++b)
#pragma unroll
        for (int c = 0; c < 3; ++c) Ss[b][c] = T(0);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        T G[3][3];
#pragma unroll
        for (int b = 0; b < 3; ++b) {
            T c[4];
#pragma unroll
            for (int a = 0; a < 4; ++a) c[a] = (a == k) ? T(3) * u[k][b] : T(4) * u[p2_edge_node(a, k)][b] - u[a][b];
            G[b][0] = c[1] - c[0];
            G[b][1] = c[2] - c[0];
            G[b][2] = c[3] - c[0];
        }
        p2_le_stress(A, idet, mu, lambda, G, Sk[k]);
#pragma unroll
        for (int b = 0; b < 3; ++b)
#pragma unroll
            for (int c = 0; c < 3; ++c) Ss[b][c] += Sk[k][b][c];
    }
    // P^(k) = ((Ss + S^(k)) / 120) A^T ; d[k][a][b] = grad lambda_a . P^(k)[b][:]
    const T s = T(1) / T(120);
    T d[4][4][3], sumd[4][3];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b) sumd[a][b] = T(0);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
#pragma unroll
        for (int b = 0; b < 3; ++b) {
            const T w0 = s * (Ss[b][0] + Sk[k][b][0]), w1 = s * (Ss[b][1] + Sk[k][b][1]), w2 = s * (Ss[b][2] + Sk[k][b][2]);
            T sum = T(0);
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const T p = A[c * 3] * w0 + A[c * 3 + 1] * w1 + A[c * 3 + 2] * w2;
                d[k][c + 1][b] = p;
                sum += p;
            }
            d[k][0][b] = -sum;
#pragma unroll
            for (int a = 0; a < 4; ++a) sumd[a][b] += d[k][a][b];
        }
    }
#pragma unroll
    for (int b = 0; b < 3; ++b) {
#pragma unroll
        for (int a = 0; a < 4; ++a) out[a][b] = T(4) * d[a][a][b] - sumd[a][b];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int c = a + 1; c < 4; ++c) out[p2_edge_node(a, c)][b] = T(4) * (d[a][c][b] + d[c][a][b]);
    }
}

// constant-strain tet with given adjugate / determinant; accumulates into out rows n[0..3]
template <typename T>
B2P2_HD void tet4_le_local_add(const T (&A)[9], const T det, const T mu, const T lambda, const T (&u)[10][3], T (&out)[10][3], const int n0,
                               const int n1, const int n2, const int n3) {
    const T idet = T(1) / det;
    T G[3][3], S[3][3];
#pragma unroll
    for (int b = 0; b < 3; ++b) {
        G[b][0] = u[n1][b] - u[n0][b];
        G[b][1] = u[n2][b] - u[n0][b];
        G[b][2] = u[n3][b] - u[n0][b];
    }
    p2_le_stress(A, idet, mu, lambda, G, S);
    const T sixth = T(1.0 / 6.0);
#pragma unroll
    for (int b = 0; b < 3; ++b) {
        const T p0 = sixth * (A[0] * S[b][0] + A[1] * S[b][1] + A[2] * S[b][2]);
        const T p1 = sixth * (A[3] * S[b][0] + A[4] * S[b][1] + A[5] * S[b][2]);
        const T p2 = sixth * (A[6] * S[b][0] + A[7] * S[b][1] + A[8] * S[b][2]);
        out[n0][b] -= p0 + p1 + p2;
        out[n1][b] += p0;
        out[n2][b] += p1;
        out[n3][b] += p2;
    }
}

// adjugate_s = B_s adjugate (rows combined with 0, +-2)
template <typename T, int S>
B2P2_HD void macro_tet4_sub_adjugate(const T (&A)[9], T (&sa)[9]) {
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            T acc = T(0);
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const int b = MacroTet4::halfB(S, r, k);
                if (b == 1) acc += A[k * 3 + c];
                if (b == -1) acc -= A[k * 3 + c];
            }
            sa[r * 3 + c] = T(2) * acc;
        }
}

template <typename T, int S>
B2P2_HD void macro_tet4_le_sub(const T (&A)[9], const T det8, const T mu, const T lambda, const T (&u)[10][3], T (&out)[10][3]) {
    T sa[9];
    macro_tet4_sub_adjugate<T, S>(A, sa);
    tet4_le_local_add(sa, det8, mu, lambda, u, out, MacroTet4::node(S, 0), MacroTet4::node(S, 1), MacroTet4::node(S, 2), MacroTet4::node(S, 3));
}

template <typename T>
B2P2_HD void macro_tet4_le_local(const T (&A)[9], const T det, const T mu, const T lambda, const T (&u)[10][3], T (&out)[10][3]) {
#pragma unroll
    for (int i = 0; i < 10; ++i)
#pragma unroll
        for (int b = 0; b < 3; ++b) out[i][b] = T(0);
    const T det8 = T(8) * det;
    macro_tet4_le_sub<T, 0>(A, det8, mu, lambda, u, out);
    macro_tet4_le_sub<T, 1>(A, det8, mu, lambda, u, out);
    macro_tet4_le_sub<T, 2>(A, det8, mu, lambda, u, out);
    macro_tet4_le_sub<T, 3>(A, det8, mu, lambda, u, out);
    macro_tet4_le_sub<T, 4>(A, det8, mu, lambda, u, out);
    macro_tet4_le_sub<T, 5>(A, det8, mu, lambda, u, out);
    macro_tet4_le_sub<T, 6>(A, det8, mu, lambda, u, out);
    macro_tet4_le_sub<T, 7>(A, det8, mu, lambda, u, out);
}

template <typename T, bool MACRO>
B2P2_HD void p2tet_le_local(const T (&A)[9], const T det, const T mu, const T lambda, const T (&u)[10][3], T (&out)[10][3]) {
    if (MACRO)
        macro_tet4_le_local(A, det, mu, lambda, u, out);
    else
        tet10_le_local(A, det, mu, lambda, u, out);
}

}  // namespace b200fem
