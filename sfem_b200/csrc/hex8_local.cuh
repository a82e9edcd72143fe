// Local (element-level) operators of the affine trilinear hex, written in the tensor "sum/difference"
// basis. Shared by the hex8 kernels and by the sshex8 stencil construction.
//
// For a 1-D linear element on [0,1] every reference matrix that the affine Q1 Laplacian / elasticity
// needs is a tensor product of
//     M  = int phi_i  phi_j  = [[1/3,1/6],[1/6,1/3]]      (mass)
//     K  = int phi_i' phi_j' = [[1,-1],[-1,1]]            (stiffness)
//     C  = int phi_i' phi_j  = [[-1/2,-1/2],[1/2,1/2]]    (derivative on the test side)
//     C^T                                                  (derivative on the trial side)
// Writing a nodal pair (u0,u1) as (s,d) = (u0+u1, u1-u0) and the result pair as (os,od) =
// (o0+o1, o1-o0) turns all four into scalings:  M: os=s/2, od=d/6 ;  K: od=2d ;  C: od=s ;  C^T: os=d.
// A 2x2x2 element vector therefore needs 24 add/sub to enter the basis, a handful of FMAs per
// geometric coefficient, and 24 add/sub (+ a factor 1/8) to leave it: ~90 flops for the Laplacian
// instead of the ~250 of the expanded closed form (hex8_laplacian_apply_fff_integral,
// operators/hex8/hex8_laplacian_inline_cpu.hpp:332-443 of the reference), same value in exact
// arithmetic.
#pragma once

namespace b200fem {

// VTK hex8 corner v -> lexicographic lattice index x + 2y + 4z
__device__ __constant__ const int kVtkToLex[8] = {0, 1, 3, 2, 4, 5, 7, 6};

// in-place forward transform along all three directions: c[x + 2y + 4z], index bit = 0 -> "s", 1 -> "d"
template <typename T>
__device__ __forceinline__ void haar_forward(T c[8]) {
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

// inverse of the output convention (os,od) -> (o0,o1) = ((os-od)/2, (os+od)/2); the 1/8 is left to the caller
template <typename T>
__device__ __forceinline__ void haar_inverse_unscaled(T c[8]) {
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

// O += sum_{j,l} t[j][l] * G_jl c   in the (s,d) basis, G_jl = int d_j phi_m d_l phi_n over the unit cube.
// t is a general (not necessarily symmetric) 3x3. Index bits: x = 1, y = 2, z = 4; bit set = "d".
template <typename T>
__device__ __forceinline__ void haar_grad_grad(const T t[3][3], const T c[8], T O[8]) {
    const T ms = T(0.5), md = T(1.0 / 6.0);
    // diagonal terms: K along j (od = 2 d_j), M along the other two
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const int bj = 1 << j;
        const int b1 = 1 << ((j + 1) % 3), b2 = 1 << ((j + 2) % 3);
#pragma unroll
        for (int a1 = 0; a1 < 2; ++a1) {
#pragma unroll
            for (int a2 = 0; a2 < 2; ++a2) {
                const int idx = bj | (a1 ? b1 : 0) | (a2 ? b2 : 0);
                const T w = T(2) * (a1 ? md : ms) * (a2 ? md : ms);
                O[idx] += (t[j][j] * w) * c[idx];
            }
        }
    }
    // mixed terms j != l: C along j (s_j -> od_j), C^T along l (d_l -> os_l), M along the third
#pragma unroll
    for (int j = 0; j < 3; ++j) {
#pragma unroll
        for (int l = 0; l < 3; ++l) {
            if (j == l) continue;
            const int k = 3 - j - l;
            const int bj = 1 << j, bl = 1 << l, bk = 1 << k;
#pragma unroll
            for (int ak = 0; ak < 2; ++ak) {
                const int in = bl | (ak ? bk : 0);   // s along j, d along l
                const int out = bj | (ak ? bk : 0);  // d along j, s along l
                O[out] += (t[j][l] * (ak ? md : ms)) * c[in];
            }
        }
    }
}

// Laplacian: t = symmetric FFF (6 entries: 00 01 02 11 12 22)
template <typename T>
__device__ __forceinline__ void hex8_laplacian_local(const T fff[6], const T u_lex[8], T out_lex[8]) {
    T c[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) c[i] = u_lex[i];
    haar_forward(c);
    const T t[3][3] = {{fff[0], fff[1], fff[2]}, {fff[1], fff[3], fff[4]}, {fff[2], fff[4], fff[5]}};
    T O[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) O[i] = T(0);
    haar_grad_grad(t, c, O);
    haar_inverse_unscaled(O);
#pragma unroll
    for (int i = 0; i < 8; ++i) out_lex[i] = T(0.125) * O[i];
}

// diagonal of the 8x8 Laplacian element matrix at lattice corner (ix,iy,iz)
template <typename T>
__device__ __forceinline__ T hex8_laplacian_diag_entry(const T fff[6], int lex) {
    const T sx = (lex & 1) ? T(1) : T(-1), sy = (lex & 2) ? T(1) : T(-1), sz = (lex & 4) ? T(1) : T(-1);
    return (fff[0] + fff[3] + fff[5]) * T(1.0 / 9.0) +
           (fff[1] * sx * sy + fff[2] * sx * sz + fff[4] * sy * sz) * T(1.0 / 6.0);
}

}  // namespace b200fem
