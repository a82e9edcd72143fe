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

// a*b + c as ONE fused instruction on the device (the compiler may not contract `acc += a*b + c*d`
// into a chain of FMAs on its own: the inner sum has to be rounded first)
B2_HD double b2_fma(double a, double b, double c) {
#ifdef __CUDA_ARCH__
    return __fma_rn(a, b, c);
#else
    return a * b + c;
#endif
}
B2_HD float b2_fma(float a, float b, float c) {
#ifdef __CUDA_ARCH__
    return __fmaf_rn(a, b, c);
#else
    return a * b + c;
#endif
}

template <typename T, int L>
struct Sshex8LaplacianRow {
    static constexpr int N = L + 1;

    T c[3][3][3];   // [dz+1][dy+1][dx+1], this row's y-class folded in (taps leaving the lattice = 0)
    T gz[3];        // z-edge term on the raw end values of the dy = 0 row: +-gz[dz+1]
    T gy[3];        // y-edge term (only on output planes 0 / L, dz = 0): +-sigma_z * gy[dy+1]
    T gx;           // x-edge term: sigma_y/2 * Syz (times sigma_z, times the 1-D mass row)
    // accumulators of three consecutive output planes. With phase P the roles are: plane zp-1 -> acc[(P+2)%3],
    // plane zp -> acc[P], plane zp+1 -> acc[(P+1)%3]. The dense-plane API below keeps P = 1 and rotates the
    // registers after every step; callers that unroll z use P = zp % 3 and never move a register.
    T acc[3][N];
    int y;

    // fff: the macro element's 6 FFF entries already widened to T (00 01 02 11 12 22)
    B2_HD void init(const T fff[6], const int y_) {
        set(fff, y_);
        clear();
    }
    B2_HD void clear() {
#pragma unroll
        for (int x = 0; x < N; ++x) acc[0][x] = acc[1][x] = acc[2][x] = T(0);
    }
    // coefficients only: the accumulators keep their values (callers that carry the last output plane of one element
    // into the first step of the next)
    B2_HD void set(const T fff[6], const int y_) {
        y = y_;
        const T h = T(1) / T(L);
        const T Sxx = fff[0] * h, Sxy = fff[1] * h, Sxz = fff[2] * h, Syy = fff[3] * h, Syz = fff[4] * h,
                Szz = fff[5] * h;
        const T m0 = T(1.0 / 6.0), m1 = T(2.0 / 3.0);
        // y-class factors: taps that leave the lattice vanish, in-plane taps of boundary rows are halved
        const T f0 = (y == 0) ? T(0) : T(1);
        const T f1 = (y == 0 || y == L) ? T(0.5) : T(1);
        const T f2 = (y == L) ? T(0) : T(1);
        const T sy = (y == 0) ? T(-1) : ((y == L) ? T(1) : T(0));
        // separable pieces (Kx = (-1,2,-1), Mx = (m0,m1,m0), Dx = (-1,0,1))
        //   even-in-x part:  A(dz,dy)[|dx|]   odd-in-x part: B(dz,dy) * dx
        const T fy[3] = {f0, f1, f2};
        const T Mt[3] = {m0, m1, m0};
        const T Kt[3] = {T(-1), T(2), T(-1)};
#pragma unroll
        for (int iz = 0; iz < 3; ++iz) {
#pragma unroll
            for (int iy = 0; iy < 3; ++iy) {
                const T mm = Mt[iy] * Mt[iz];
                const T km = Syy * (Kt[iy] * Mt[iz]) + Szz * (Mt[iy] * Kt[iz]);
                // dy dz term (zero unless both are +-1): -1/2 Syz dy dz M(dx)
                const T dydz = (iy == 1 || iz == 1) ? T(0) : ((iy == iz) ? T(-0.5) : T(0.5)) * Syz;
                // dx-odd term: -1/2 (Sxy dy M(dz) + Sxz dz M(dy)) * dx
                const T oy = (iy == 1) ? T(0) : ((iy == 2) ? T(-0.5) : T(0.5)) * Sxy * Mt[iz];
                const T oz = (iz == 1) ? T(0) : ((iz == 2) ? T(-0.5) : T(0.5)) * Sxz * Mt[iy];
                const T even_side = -Sxx * mm + m0 * km + m0 * dydz;    // dx = +-1
                const T even_mid = T(2) * Sxx * mm + m1 * km + m1 * dydz;  // dx = 0
                const T odd = oy + oz;                                   // coefficient of dx
                c[iz][iy][0] = (even_side - odd) * fy[iy];
                c[iz][iy][1] = even_mid * fy[iy];
                c[iz][iy][2] = (even_side + odd) * fy[iy];
            }
        }
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            gz[d] = T(0.5) * Sxy * sy * Mt[d];    // times sigma_x at use
            gy[d] = T(0.5) * Sxz * Mt[d] * fy[d];  // times sigma_x sigma_z at use
        }
        gx = T(0.5) * Syz * sy;  // times sigma_z at use
    }

    // acc[x] += sum_dx cc[dx+1] * r[x+dx] with the x-boundary rule (h0 = r[0]/2, hL = r[L]/2)
    B2_HD static void row_taps(const T cc[3], const T r[N], const T h0, const T hL, T acc[N]) {
        acc[0] = b2_fma(cc[1], h0, b2_fma(cc[2], r[1], acc[0]));
#pragma unroll
        for (int x = 1; x < L; ++x) acc[x] = b2_fma(cc[0], r[x - 1], b2_fma(cc[1], r[x], b2_fma(cc[2], r[x + 1], acc[x])));
        acc[L] = b2_fma(cc[0], r[L - 1], b2_fma(cc[1], hL, acc[L]));
    }

    // Consume input plane zp. ZC = 0: zp == 0, 1: 0 < zp < L, 2: zp == L (compile-time plane class, so
    // the interior planes carry no boundary code). `load_row(iy, r)` fills r[0..L] with the lattice values of
    // row clamp(y + iy - 1, 0, L) of plane zp (clamped rows carry a zero coefficient). Reads only.
    template <int ZC, int P, class RowLoader>
    B2_HD void step_class_with(RowLoader &&load_row) {
        constexpr bool zlo = (ZC == 0), zhi = (ZC == 2);
        T(&accP)[N] = acc[(P + 2) % 3];
        T(&accC)[N] = acc[P % 3];
        T(&accN)[N] = acc[(P + 1) % 3];
        if (zhi) {
#pragma unroll
            for (int x = 0; x < N; ++x) accC[x] *= T(2);  // exact; undone below => dz = 0 taps get 1/2
        }
#pragma unroll
        for (int iy = 0; iy < 3; ++iy) {
            T r[N];
            load_row(iy, r);
            const T h0 = T(0.5) * r[0], hL = T(0.5) * r[L];
            if (!zlo) row_taps(c[2][iy], r, h0, hL, accP);  // output plane zp-1 sees this plane at dz = +1
            row_taps(c[1][iy], r, h0, hL, accC);
            if (!zhi) row_taps(c[0][iy], r, h0, hL, accN);  // output plane zp+1 sees it at dz = -1
            if (iy == 1) {
                // edges parallel to z: raw end values of the row itself
                if (!zlo) {
                    accP[0] = b2_fma(-gz[2], r[0], accP[0]);
                    accP[L] = b2_fma(gz[2], r[L], accP[L]);
                }
                accC[0] = b2_fma(-gz[1], r[0], accC[0]);
                accC[L] = b2_fma(gz[1], r[L], accC[L]);
                if (!zhi) {
                    accN[0] = b2_fma(-gz[0], r[0], accN[0]);
                    accN[L] = b2_fma(gz[0], r[L], accN[L]);
                }
            }
            if (zlo || zhi) {
                // edges parallel to y (x and z extreme); applied after the halving below, so pre-double
                const T sz2 = zlo ? T(-2) : T(2);
                accC[0] = b2_fma(-sz2 * gy[iy], r[0], accC[0]);
                accC[L] = b2_fma(sz2 * gy[iy], r[L], accC[L]);
                if (iy == 1 && (y == 0 || y == L)) {
                    // edges parallel to x (y and z extreme): 1-D mass row along x (pre-doubled as well)
                    const T g = gx * sz2;
                    const T m[3] = {g * T(1.0 / 6.0), g * T(2.0 / 3.0), g * T(1.0 / 6.0)};
                    row_taps(m, r, h0, hL, accC);
                }
            }
        }
        if (zlo || zhi) {
#pragma unroll
            for (int x = 0; x < N; ++x) accC[x] *= T(0.5);
        }
    }

    // The edge terms of a boundary plane (z = 0: sgn2 = -2, z = L: sgn2 = +2) applied in one piece, for callers that run the
    // plane through the interior-class code (accC pre-doubled for z = L, halved afterwards for both): r0 / rL = the end values
    // of the three rows, rmid = the dy = 0 row. gx is zero on lanes whose row is not on a y-boundary, so the x-edge taps run
    // unconditionally.
    B2_HD void z_boundary_terms(T (&accC)[N], const T sgn2, const T (&r0)[3], const T (&rL)[3], const T (&rmid)[N]) const {
#pragma unroll
        for (int iy = 0; iy < 3; ++iy) {
            accC[0] = b2_fma(-sgn2 * gy[iy], r0[iy], accC[0]);
            accC[L] = b2_fma(sgn2 * gy[iy], rL[iy], accC[L]);
        }
        const T g = gx * sgn2;
        const T m[3] = {g * T(1.0 / 6.0), g * T(2.0 / 3.0), g * T(1.0 / 6.0)};
        row_taps(m, rmid, T(0.5) * rmid[0], T(0.5) * rmid[L], accC);
    }

    // dense-plane form: `plane` points at this element's plane zp; row yy starts at plane[yy * ystride]
    template <int ZC>
    B2_HD void step_class(const T *plane, const int ystride) {
        const int yrow = y;
        step_class_with<ZC, 1>([plane, ystride, yrow](const int iy, T(&r)[N]) {
            int yy = yrow + iy - 1;
            yy = yy < 0 ? 0 : (yy > L ? L : yy);
            const T *row = plane + yy * ystride;
#pragma unroll
            for (int x = 0; x < N; ++x) r[x] = row[x];
        });
    }

    B2_HD void step(const int zp, const T *plane, const int ystride) {
        if (zp == 0)
            step_class<0>(plane, ystride);
        else if (zp == L)
            step_class<2>(plane, ystride);
        else
            step_class<1>(plane, ystride);
    }

    // rolled-loop form with a caller-supplied row loader (any shared-memory layout); pair with rotate()
    template <class RowLoader>
    B2_HD void step_with(const int zp, RowLoader &&load_row) {
        if (zp == 0)
            step_class_with<0, 1>(load_row);
        else if (zp == L)
            step_class_with<2, 1>(load_row);
        else
            step_class_with<1, 1>(load_row);
    }

    // compile-time plane index: phase zp % 3, no register rotation (see `acc`)
    template <int ZP, class RowLoader>
    B2_HD void step_static(RowLoader &&load_row) {
        step_class_with<(ZP == 0 ? 0 : (ZP == L ? 2 : 1)), ZP % 3>(load_row);
    }
    // after step_static<ZP>: the finished output plane ZP-1 (ZP >= 1) ...
    template <int ZP>
    B2_HD T (&finished_plane())[N] {
        return acc[(ZP + 2) % 3];
    }
    // ... which then becomes the accumulator of plane ZP+2
    template <int ZP>
    B2_HD void recycle() {
#pragma unroll
        for (int x = 0; x < N; ++x) acc[(ZP + 2) % 3][x] = T(0);
    }

    // After step(zp) and a barrier: emit finished output plane zp-1 (zp >= 1), then rotate.
    B2_HD void write_and_rotate(const int zp, T *out_plane_prev, const int ystride) {
        if (zp > 0) {
            T *row = out_plane_prev + y * ystride;
#pragma unroll
            for (int x = 0; x < N; ++x) row[x] = acc[0][x];
        }
        rotate();
    }

    B2_HD void rotate() {
#pragma unroll
        for (int x = 0; x < N; ++x) {
            acc[0][x] = acc[1][x];
            acc[1][x] = acc[2][x];
            acc[2][x] = T(0);
        }
    }

    // After the last step/rotate: acc[0] holds output plane L.
    B2_HD void write_last(T *out_plane_L, const int ystride) {
        T *row = out_plane_L + y * ystride;
#pragma unroll
        for (int x = 0; x < N; ++x) row[x] = acc[0][x];
    }
};

}  // namespace b200fem
