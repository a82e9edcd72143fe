// Mass (consistent: y += M x) and lumped mass of the isoparametric Q1 hex, 27-point Gauss on [0,1]^3: hex8_apply_mass /
// hex8_assemble_lumped_mass (operators/hex8/hex8_mass.cpp:8-103; point kernels hex8_mass_apply_points / hex8_lumped_mass_points,
// hex8_mass_inline_cpu.hpp:135-286; points and weights q27_*, hex8_quadrature.hpp:46-84 -- the reference's 15-digit literals).
// Per point: out_i += w det J(q) phi_i(q) sum_j phi_j(q) u_j (lumped: the row sum, sum_j phi_j = 1).
// Corners in lexicographic order (x + 2 y + 4 z). Shared by from_points.cu (hex8) and sshex8_extras.cu (macro element of sshex8).
#pragma once
#include "hex8_le_core.cuh"

namespace b200fem {

template <typename T>
__device__ __forceinline__ void hex8_mass_local(const T (&X)[3][8], const T (&U)[8], T (&O)[8]) {
    const T qp[3] = {T(0.112701665379258), T(0.5), T(0.887298334620742)};
    // weight by the number of centre coordinates among (qx, qy, qz)
    const T qw[4] = {T(0.021433470507545), T(0.034293552812071), T(0.054869684499314), T(0.087791495198903)};
#pragma unroll
    for (int i = 0; i < 8; ++i) O[i] = T(0);
#pragma unroll 1
    for (int q = 0; q < 27; ++q) {
        const int ix = q % 3, iy = (q / 3) % 3, iz = q / 9;
        const T x = qp[ix], yq = qp[iy], z = qp[iz];
        const T w = qw[(ix == 1) + (iy == 1) + (iz == 1)];
        T phi[8], J[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) J[k] = T(0);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int bx = i & 1, by = (i >> 1) & 1, bz = (i >> 2) & 1;
            const T sx = bx ? x : T(1) - x, sy = by ? yq : T(1) - yq, sz = bz ? z : T(1) - z;
            phi[i] = sx * sy * sz;
            const T gx = (bx ? T(1) : T(-1)) * sy * sz, gy = sx * (by ? T(1) : T(-1)) * sz, gz = sx * sy * (bz ? T(1) : T(-1));
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                J[r * 3 + 0] = le_fma(X[r][i], gx, J[r * 3 + 0]);
                J[r * 3 + 1] = le_fma(X[r][i], gy, J[r * 3 + 1]);
                J[r * 3 + 2] = le_fma(X[r][i], gz, J[r * 3 + 2]);
            }
        }
        const T det = J[0] * (J[4] * J[8] - J[5] * J[7]) - J[1] * (J[3] * J[8] - J[5] * J[6]) + J[2] * (J[3] * J[7] - J[4] * J[6]);
        T uq = T(0);
#pragma unroll
        for (int i = 0; i < 8; ++i) uq = le_fma(phi[i], U[i], uq);
        const T val = w * det * uq;
#pragma unroll
        for (int i = 0; i < 8; ++i) O[i] = le_fma(phi[i], val, O[i]);
    }
}

}  // namespace b200fem
