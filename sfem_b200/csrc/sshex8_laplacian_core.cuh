// Per-thread core of the sshex8 (semi-structured hex8) Laplacian: one thread owns one x-row
// (fixed local y) of a macro element's (L+1)^3 lattice and marches over z.
//
// What it computes is the reference's affine_sshex8_laplacian_stencil_apply_fff
// (operators/sshex8/sshex8_laplacian.cpp:698-769): v = A_loc * eu, A_loc = the assembly of ONE 8x8
// micro-element matrix (hex8_laplacian_matrix_fff_integral of fff*h, h = 1/L) over the L^3 micro
// elements. The reference evaluates it as a 27-point interior stencil plus 26 generated boundary-class
// stencils (operators/stencil/stencil3.hpp:167-184, sshex8_skeleton_stencil.hpp:1683-1722).
//
// Restatement used here (exact in exact arithmetic, derived from the tensor structure of the affine
// Q1 Laplacian -- see DESIGN.md "sshex8 Laplacian"): with the 1-D assembled taps
//     M = (1/6, 2/3, 1/6),  K = (-1, 2, -1),  D(d) = d          (d = -1, 0, +1)
// and S = fff * h (symmetric 3x3), the interior stencil is
//     S27[dz][dy][dx] = Sxx K(dx)M(dy)M(dz) + Syy M(dx)K(dy)M(dz) + Szz M(dx)M(dy)K(dz)
//                       - 1/2 ( Sxy dx dy M(dz) + Sxz dx dz M(dy) + Syz dy dz M(dx) ).
// For a node on the macro-element boundary the assembled row is the SAME stencil restricted to the
// lattice, with every tap that stays in the node's own boundary plane (d = 0 in a direction in which
// the node is on the boundary) multiplied by 1/2, plus, on the 12 macro edges only,
//     + 1/2 S_jl sigma_j sigma_l (M_k u)(p)         (j,l = the two directions the edge is extreme in,
//                                                    sigma = -1 on the low side, +1 on the high side,
//                                                    M_k = the 1-D mass taps along the edge, centre
//                                                    tap halved at the edge's end points).
// This needs ONE coefficient set per element instead of 27 and ~1.08x the minimal flop count.
//
// The struct below is the register state of one thread. It is __host__ __device__ so that
// tests/test_sshex8_core.py can run exactly this code on the CPU against the oracle.
#pragma once

#ifdef __CUDACC__
#define B2_HD __host__ __device__ __forceinline__
#else
#define B2_HD inline
#endif

namespace b200fem {

template <typename T, int L>
struct Sshex8LaplacianRow {
    static constexpr int N = L + 1;

    T c[3][3][3];   // [dz+1][dy+1][dx+1], this row's y-class folded in (taps leaving the lattice = 0)
    T gz[3];        // z-edge term on the raw end values of the dy = 0 row: +-gz[dz+1]
    T gy[3];        // y-edge term (only on output planes 0 / L, dz = 0): +-sigma_z * gy[dy+1]
    T gx;           // x-edge term: sigma_y/2 * Syz (times sigma_z, times the 1-D mass row)
    T accP[N], accC[N], accN[N];  // output planes zp-1, zp, zp+1
    int y;

    // fff: the macro element's 6 FFF entries already widened to T (00 01 02 11 12 22)
    B2_HD void init(const T fff[6], const int y_) {
        y = y_;
        const T h = T(1) / T(L);
        const T Sxx = fff[0] * h, Sxy = fff[1] * h, Sxz = fff[2] * h, Syy = fff[3] * h, Syz = fff[4] * h,
                Szz = fff[5] * h;
        const T Mt[3] = {T(1.0 / 6.0), T(2.0 / 3.0), T(1.0 / 6.0)};
        const T Kt[3] = {T(-1), T(2), T(-1)};
        // y-class factors: taps that leave the lattice vanish, in-plane taps of boundary rows are halved
        T fy[3] = {T(1), T(1), T(1)};
        T sy = T(0);
        if (y == 0) {
            fy[0] = T(0);
            fy[1] = T(0.5);
            sy = T(-1);
        } else if (y == L) {
            fy[2] = T(0);
            fy[1] = T(0.5);
            sy = T(1);
        }
#pragma unroll
        for (int iz = 0; iz < 3; ++iz) {
#pragma unroll
            for (int iy = 0; iy < 3; ++iy) {
#pragma unroll
                for (int ix = 0; ix < 3; ++ix) {
                    const T dx = T(ix - 1), dy = T(iy - 1), dz = T(iz - 1);
                    const T v = Sxx * Kt[ix] * Mt[iy] * Mt[iz] + Syy * Mt[ix] * Kt[iy] * Mt[iz] +
                                Szz * Mt[ix] * Mt[iy] * Kt[iz] -
                                T(0.5) * (Sxy * dx * dy * Mt[iz] + Sxz * dx * dz * Mt[iy] + Syz * dy * dz * Mt[ix]);
                    c[iz][iy][ix] = v * fy[iy];
                }
            }
        }
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            gz[d] = T(0.5) * Sxy * sy * Mt[d];   // times sigma_x at use
            gy[d] = T(0.5) * Sxz * Mt[d] * fy[d]; // times sigma_x sigma_z at use
        }
        gx = T(0.5) * Syz * sy;  // times sigma_z at use
#pragma unroll
        for (int x = 0; x < N; ++x) accP[x] = accC[x] = accN[x] = T(0);
    }

    // acc[x] += sum_dx cc[dx+1] * r[x+dx] with the x-boundary rule (h0 = r[0]/2, hL = r[L]/2)
    B2_HD static void row_taps(const T cc[3], const T r[N], const T h0, const T hL, T acc[N]) {
        acc[0] += cc[1] * h0 + cc[2] * r[1];
#pragma unroll
        for (int x = 1; x < L; ++x) acc[x] += cc[0] * r[x - 1] + cc[1] * r[x] + cc[2] * r[x + 1];
        acc[L] += cc[0] * r[L - 1] + cc[1] * hL;
    }

    // Consume input plane zp. `plane` points at this element's plane zp; row yy starts at
    // plane[yy * ystride]. Reads only; the caller synchronises before write_plane().
    B2_HD void step(const int zp, const T *plane, const int ystride) {
        const bool zlo = (zp == 0), zhi = (zp == L);
        if (zhi) {
#pragma unroll
            for (int x = 0; x < N; ++x) accC[x] *= T(2);  // exact; undone below => dz = 0 taps get 1/2
        }
#pragma unroll
        for (int iy = 0; iy < 3; ++iy) {
            int yy = y + iy - 1;
            yy = yy < 0 ? 0 : (yy > L ? L : yy);  // clamped rows carry a zero coefficient
            const T *row = plane + yy * ystride;
            T r[N];
#pragma unroll
            for (int x = 0; x < N; ++x) r[x] = row[x];
            const T h0 = T(0.5) * r[0], hL = T(0.5) * r[L];
            if (!zlo) row_taps(c[2][iy], r, h0, hL, accP);  // output plane zp-1 sees this plane at dz = +1
            row_taps(c[1][iy], r, h0, hL, accC);
            if (!zhi) row_taps(c[0][iy], r, h0, hL, accN);  // output plane zp+1 sees it at dz = -1
            if (iy == 1) {
                // edges parallel to z: raw end values of the row itself
                if (!zlo) {
                    accP[0] -= gz[2] * r[0];
                    accP[L] += gz[2] * r[L];
                }
                accC[0] -= gz[1] * r[0];
                accC[L] += gz[1] * r[L];
                if (!zhi) {
                    accN[0] -= gz[0] * r[0];
                    accN[L] += gz[0] * r[L];
                }
            }
        }
        if (zlo || zhi) {
#pragma unroll
            for (int x = 0; x < N; ++x) accC[x] *= T(0.5);
            // terms that live on the z = 0 / z = L planes only and are not halved
            const T sz = zlo ? T(-1) : T(1);
#pragma unroll
            for (int iy = 0; iy < 3; ++iy) {
                int yy = y + iy - 1;
                yy = yy < 0 ? 0 : (yy > L ? L : yy);
                const T *row = plane + yy * ystride;
                // edges parallel to y (x and z extreme)
                accC[0] -= sz * gy[iy] * row[0];
                accC[L] += sz * gy[iy] * row[L];
            }
            if (y == 0 || y == L) {
                // edges parallel to x (y and z extreme): 1-D mass row along x
                const T *row = plane + y * ystride;
                T r[N];
#pragma unroll
                for (int x = 0; x < N; ++x) r[x] = row[x];
                const T g = gx * sz;
                const T m[3] = {g * T(1.0 / 6.0), g * T(2.0 / 3.0), g * T(1.0 / 6.0)};
                row_taps(m, r, T(0.5) * r[0], T(0.5) * r[L], accC);
            }
        }
    }

    // After step(zp) and a barrier: emit finished output plane zp-1 (zp >= 1), then rotate.
    B2_HD void write_and_rotate(const int zp, T *out_plane_prev, const int ystride) {
        if (zp > 0) {
            T *row = out_plane_prev + y * ystride;
#pragma unroll
            for (int x = 0; x < N; ++x) row[x] = accP[x];
        }
#pragma unroll
        for (int x = 0; x < N; ++x) {
            accP[x] = accC[x];
            accC[x] = accN[x];
            accN[x] = T(0);
        }
    }

    // After the last step/rotate: accP holds output plane L.
    B2_HD void write_last(T *out_plane_L, const int ystride) {
        T *row = out_plane_L + y * ystride;
#pragma unroll
        for (int x = 0; x < N; ++x) row[x] = accP[x];
    }
};

}  // namespace b200fem
