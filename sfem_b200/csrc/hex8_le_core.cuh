// Local linear-elasticity operator of ONE affine trilinear hex (a micro element of a semi-structured
// macro element), evaluated from "gradient moments" in the tensor sum/difference (Haar) basis.
//
// Replaces the reference's per-micro-element 24x24 mat-vec `out += K^T u` with
// K = hex8_linear_elasticity_matrix(mu, lambda, sub_adjugate, sub_determinant)
// (operators/sshex8/sshex8_linear_elasticity.cpp:360-389; matrix
// operators/hex8/hex8_linear_elasticity_inline_cpu.hpp:207-2456): 576 FMA and a 576-entry matrix per
// micro element become ~300 FMA + 144 add/sub with 9 + 6 per-macro-element constants in registers.
//
// Derivation. With (s, d) = (u0 + u1, u1 - u0) per direction, a trilinear component is
//     u_b(xi) = sum_idx c_b[idx] prod_dir (bit(idx, dir) ? tau_dir : 1/2),   tau = xi - 1/2,
// so the reference gradient d_j u_b only has the monomials {1, tau_k, tau_l, tau_k tau_l} (k, l != j)
// with coefficients c_b[idx | bit j] (times 1/4, 1/2, 1/2, 1). For an affine element the maps
//     F' = g A  (g = reference gradient, A = adjugate; F = F'/det is the physical gradient, cf.
//               frontend/ops/generated/.../linear_elasticity_d3_tensor_product_local.hpp:264-292),
//     sigma' = lambda tr(F') I + mu (F' + F'^T),       P' = sigma' A^T
// are constant over the element, hence act monomial by monomial; testing with d_j of the Haar test
// functions (which do not depend on tau_j) keeps, per monomial, only the j's that produced it. The seven
// monomial groups and their combined input x output x quadrature scales are
//     1            : j in {x,y,z}, scale 1/2      tau_k : the two j != k, scale 1/6
//     tau_k tau_l  : the remaining j, scale 1/18
// (the same numbers as haar_grad_grad in hex8_local.cuh). The scales, 1/det and the 1/8 of the inverse
// transform are folded into three (lambda_g, mu_g) pairs per macro element.
//
// __host__ __device__: tests/test_sshex8_le_core.py runs exactly this code on the CPU against the oracle.
#pragma once

#ifdef __CUDACC__
#define B2LE_HD __host__ __device__ __forceinline__
#else
#define B2LE_HD inline
#endif

namespace b200fem {

// a * b + c as ONE instruction on the device. The compiler contracts `x * y + z` on its own, but not
// `acc += x0 * y0 + x1 * y1 + x2 * y2` (the inner sum has to be rounded first): written as a chain of these the
// accumulate costs 3 FMAs instead of MUL + 2 FMA + ADD.
B2LE_HD double le_fma(double a, double b, double c) {
#ifdef __CUDA_ARCH__
    return __fma_rn(a, b, c);
#else
    return a * b + c;
#endif
}
B2LE_HD float le_fma(float a, float b, float c) {
#ifdef __CUDA_ARCH__
    return __fmaf_rn(a, b, c);
#else
    return a * b + c;
#endif
}

template <typename T>
struct Hex8LEConsts {
    T A[9];     // adjugate of the (micro) element, row-major A[j*3+k]
    T lam[3];   // lambda * scale_g / (8 det), g = 0: monomial 1, 1: tau, 2: tau tau
    T mu[3];    // mu * scale_g / (8 det)
    T mu2[3];   // 2 mu_g

    // macro adjugate / determinant and h = 1/L: sub_adjugate = A h^2, sub_determinant = det h^3
    // (hex8_sub_adj_0, operators/sshex8/sshex8.hpp:26-42)
    B2LE_HD void init(const T macro_adjugate[9], const T macro_det, const T h, const T mu_, const T lambda_) {
        const T h2 = h * h;
#pragma unroll
        for (int k = 0; k < 9; ++k) A[k] = macro_adjugate[k] * h2;
        const T det = macro_det * (h2 * h);
        const T base = T(0.125) / det;
        const T sc[3] = {T(0.5), T(1.0 / 6.0), T(1.0 / 18.0)};
#pragma unroll
        for (int g = 0; g < 3; ++g) {
            lam[g] = lambda_ * sc[g] * base;
            mu[g] = mu_ * sc[g] * base;
            mu2[g] = T(2) * mu[g];
        }
    }
};

template <typename T>
B2LE_HD void haar3_forward(T c[8]) {
#pragma unroll
    for (int bit = 1; bit < 8; bit <<= 1) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (!(i & bit)) {
                const T a = c[i], b = c[i | bit];
                c[i] = a + b;
                c[i | bit] = b - a;
            }
        }
    }
}

template <typename T>
B2LE_HD void haar3_inverse_unscaled(T c[8]) {
#pragma unroll
    for (int bit = 1; bit < 8; bit <<= 1) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (!(i & bit)) {
                const T s = c[i], d = c[i | bit];
                c[i] = s - d;
                c[i | bit] = s + d;
            }
        }
    }
}

// One monomial group. J0/J1/J2 = whether direction j takes part; in[b][j] = gradient coefficient of
// component b along j (ignored when !Jj); the result is ACCUMULATED into out[a][j].
template <typename T, bool J0, bool J1, bool J2>
B2LE_HD void le_group(const T A[9], const T lam, const T mu, const T mu2, const T in[3][3], T out[3][3]) {
    constexpr bool J[3] = {J0, J1, J2};
    T F[3][3];  // F'[b][k] = sum_j in[b][j] A[j][k]
#pragma unroll
    for (int b = 0; b < 3; ++b) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            T v = T(0);
            bool first = true;
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                if (!J[j]) continue;
                v = first ? in[b][j] * A[j * 3 + k] : le_fma(in[b][j], A[j * 3 + k], v);
                first = false;
            }
            F[b][k] = v;
        }
    }
    const T ltr = lam * (F[0][0] + F[1][1] + F[2][2]);
    T S[3][3];
#pragma unroll
    for (int a = 0; a < 3; ++a) S[a][a] = le_fma(mu2, F[a][a], ltr);
    S[0][1] = S[1][0] = mu * (F[0][1] + F[1][0]);
    S[0][2] = S[2][0] = mu * (F[0][2] + F[2][0]);
    S[1][2] = S[2][1] = mu * (F[1][2] + F[2][1]);
#pragma unroll
    for (int a = 0; a < 3; ++a) {
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            if (!J[j]) continue;
            out[a][j] = le_fma(S[a][2], A[j * 3 + 2], le_fma(S[a][1], A[j * 3 + 1], le_fma(S[a][0], A[j * 3 + 0], out[a][j])));
        }
    }
}

// c[b][idx]: Haar coefficients of the three displacement components (lexicographic corner order
// x + 2y + 4z before the transform). O[a][idx] receives the element vector in the (os, od) basis, to be
// passed through haar3_inverse_unscaled (all scales are already folded into K). O is OVERWRITTEN.
template <typename T>
B2LE_HD void hex8_le_haar_apply(const Hex8LEConsts<T> &K, const T c[3][8], T O[3][8]) {
#pragma unroll
    for (int a = 0; a < 3; ++a) O[a][0] = T(0);
    T in[3][3], out[3][3];
    // monomial 1: g_bx = c[1], g_by = c[2], g_bz = c[4] -> O[1], O[2], O[4]
#pragma unroll
    for (int b = 0; b < 3; ++b) {
        in[b][0] = c[b][1]; in[b][1] = c[b][2]; in[b][2] = c[b][4];
        out[b][0] = out[b][1] = out[b][2] = T(0);
    }
    le_group<T, true, true, true>(K.A, K.lam[0], K.mu[0], K.mu2[0], in, out);
#pragma unroll
    for (int a = 0; a < 3; ++a) { O[a][1] = out[a][0]; O[a][2] = out[a][1]; O[a][4] = out[a][2]; }

    // tau_x: j in {y, z}: g_by = c[3], g_bz = c[5] -> O[3], O[5]
    T o3[3], o5[3], o6[3];
#pragma unroll
    for (int b = 0; b < 3; ++b) {
        in[b][1] = c[b][3]; in[b][2] = c[b][5];
        out[b][1] = out[b][2] = T(0);
    }
    le_group<T, false, true, true>(K.A, K.lam[1], K.mu[1], K.mu2[1], in, out);
#pragma unroll
    for (int a = 0; a < 3; ++a) { o3[a] = out[a][1]; o5[a] = out[a][2]; }
    // tau_y: j in {x, z}: g_bx = c[3], g_bz = c[6] -> O[3], O[6]
#pragma unroll
    for (int b = 0; b < 3; ++b) {
        in[b][0] = c[b][3]; in[b][2] = c[b][6];
        out[b][0] = o3[b]; out[b][2] = T(0);
    }
    le_group<T, true, false, true>(K.A, K.lam[1], K.mu[1], K.mu2[1], in, out);
#pragma unroll
    for (int a = 0; a < 3; ++a) { o3[a] = out[a][0]; o6[a] = out[a][2]; }
    // tau_z: j in {x, y}: g_bx = c[5], g_by = c[6] -> O[5], O[6]
#pragma unroll
    for (int b = 0; b < 3; ++b) {
        in[b][0] = c[b][5]; in[b][1] = c[b][6];
        out[b][0] = o5[b]; out[b][1] = o6[b];
    }
    le_group<T, true, true, false>(K.A, K.lam[1], K.mu[1], K.mu2[1], in, out);
#pragma unroll
    for (int a = 0; a < 3; ++a) { O[a][3] = o3[a]; O[a][5] = out[a][0]; O[a][6] = out[a][1]; }

    // tau tau: one j each, all three read c[7] and feed O[7]
    T o7[3] = {T(0), T(0), T(0)};
#pragma unroll
    for (int b = 0; b < 3; ++b) { in[b][0] = c[b][7]; out[b][0] = o7[b]; }
    le_group<T, true, false, false>(K.A, K.lam[2], K.mu[2], K.mu2[2], in, out);
#pragma unroll
    for (int b = 0; b < 3; ++b) { in[b][1] = c[b][7]; out[b][1] = out[b][0]; }
    le_group<T, false, true, false>(K.A, K.lam[2], K.mu[2], K.mu2[2], in, out);
#pragma unroll
    for (int b = 0; b < 3; ++b) { in[b][2] = c[b][7]; out[b][2] = out[b][1]; }
    le_group<T, false, false, true>(K.A, K.lam[2], K.mu[2], K.mu2[2], in, out);
#pragma unroll
    for (int a = 0; a < 3; ++a) O[a][7] = out[a][2];
}

// Nodal form: u[b][lex] -> out[a][lex] (OVERWRITTEN), lex = x + 2y + 4z.
template <typename T>
B2LE_HD void hex8_le_local_apply(const Hex8LEConsts<T> &K, const T u[3][8], T out[3][8]) {
    T c[3][8];
#pragma unroll
    for (int b = 0; b < 3; ++b) {
#pragma unroll
        for (int i = 0; i < 8; ++i) c[b][i] = u[b][i];
        haar3_forward(c[b]);
    }
    hex8_le_haar_apply(K, c, out);
#pragma unroll
    for (int a = 0; a < 3; ++a) haar3_inverse_unscaled(out[a]);
}

// Assembled diagonal blocks at a lattice node of a macro element. Entry (a, b) of the 24x24 micro
// matrix at corner (s_x, s_y, s_z) in {-1,+1}^3 is sum_jl T^{ab}_{jl} G_jl[m][m] with G_jj[m][m] = 1/9,
// G_jl[m][m] = s_j s_l / 12 and T^{ab}_{jl} = [lambda A_ja A_lb + mu A_jb A_la + mu delta_ab (A A^T)_jl]/det.
// Summed over the micro elements around lattice node (x, y, z): n_x n_y n_z / 9 on the diagonal part and
// (sum s_j)(sum s_l) n_k / 12 on the mixed part, n = 1 on a macro face, 2 inside; sum s = -1 / 0 / +1 on the
// low face / inside / high face. Replaces affine_sshex8_linear_elasticity_diag / _block_diag_sym
// (operators/sshex8/sshex8_linear_elasticity.cpp:971-1256). `blk` = 6 entries (00 01 02 11 12 22).
template <typename T>
B2LE_HD void sshex8_le_node_block(const T A[9] /* sub adjugate */, const T det /* sub determinant */, const T mu,
                                  const T lambda, const int L, const int x, const int y, const int z, T blk[6]) {
    const int p[3] = {x, y, z};
    T n[3], s[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        n[d] = (p[d] == 0 || p[d] == L) ? T(1) : T(2);
        s[d] = (p[d] == 0) ? T(-1) : (p[d] == L ? T(1) : T(0));
    }
    T w[3][3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
#pragma unroll
        for (int l = 0; l < 3; ++l) {
            if (j == l) {
                w[j][l] = n[0] * n[1] * n[2] * T(1.0 / 9.0);
            } else {
                w[j][l] = s[j] * s[l] * n[3 - j - l] * T(1.0 / 12.0);
            }
        }
    }
    const T idet = T(1) / det;
    // sum_jl w_jl (A A^T)_jl
    T aat = T(0);
#pragma unroll
    for (int j = 0; j < 3; ++j) {
#pragma unroll
        for (int l = 0; l < 3; ++l)
            aat += w[j][l] * (A[j * 3 + 0] * A[l * 3 + 0] + A[j * 3 + 1] * A[l * 3 + 1] + A[j * 3 + 2] * A[l * 3 + 2]);
    }
    int slot = 0;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
#pragma unroll
        for (int b = a; b < 3; ++b) {
            T v = T(0);
#pragma unroll
            for (int j = 0; j < 3; ++j) {
#pragma unroll
                for (int l = 0; l < 3; ++l)
                    v += w[j][l] * (lambda * A[j * 3 + a] * A[l * 3 + b] + mu * A[j * 3 + b] * A[l * 3 + a]);
            }
            if (a == b) v += mu * aat;
            blk[slot++] = v * idet;
        }
    }
}

}  // namespace b200fem
