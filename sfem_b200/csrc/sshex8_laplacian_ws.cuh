// sshex8 Laplacian, warp-specialised plane-streaming kernels: compute warps only read shared memory and
// issue FMAs, mover warps own every global access. Two kernels live here: sshex8_laplacian_ws_kernel (round 1: four hand-offs
// per plane; kept for A/B) and sshex8_laplacian_ws2_kernel further down (round 2, the default: two hand-offs per plane,
// static accumulator phases, optional in-kernel wait for the ghost block of a distributed apply). What follows describes what
// they share. Same arithmetic (sshex8_laplacian_core.cuh), same element
// headers and same plane ring as sshex8_laplacian_stream.cuh; what changes is who does what, and how the interior
// of a plane travels.
//
// Why (ncu of the one-warp-does-everything kernel, profiles/r01e_ncu_sshex8_stream.details.txt):
//   * 2 warps per scheduler at 250 registers, each alternating between a DFMA stream and the integer / LSU chains of
//     its own copies: issue slots 41 % busy, FP64 pipe 37 % busy, 1.6 cycles of every 4.8 between two instructions of
//     a warp spent in fixed-latency waits nobody covers;
//   * the L1 data pipe is the busiest unit (71 %): an 8-byte LDGSTS costs one shared-memory wavefront per 32-byte
//     SECTOR it fetches (52.5 M wavefronts for 191 M nodes, 4x what the bytes need), and the shared loads of the
//     arithmetic pay 2.7-way bank conflicts between the element slots of a warp (layout [slot][y][x], 81 doubles
//     per slot-plane).
// Here
//   * a CTA is NC compute warps + NC mover warps (one mover per compute warp; setmaxnreg moves the registers to the
//     compute warps). Hand-off through mbarriers in shared memory, per compute warp: full[D] / empty[D] for the ring
//     of input planes, sfull[S] / sempty[S] for the staged output planes. No CTA-wide barrier after the prologue;
//   * for compact numberings (interior block of an element = one contiguous id run, x fastest) the (L-1)^2 interior
//     values of a plane arrive as ONE 1-D tensor-map TMA copy per element (box of BOX elements starting at the
//     16-byte boundary below the run; the over-fetch is at most 3 neighbours that sit in the same sectors anyway;
//     out-of-range coordinates are zero-filled by the hardware). Only the 4 L skeleton
//     positions of a plane (and the whole planes z = 0 and z = L) travel per node;
//   * a slot-plane buffer is [interior run | skeleton: row y=0, row y=L, column x=0, column x=L] with per-slot offsets
//     (WsLayout) chosen by exhaustive search so that the 27 shared loads of a z-step are nearly conflict-free.
#pragma once
#include <cuda.h>

#include <type_traits>

#include "common.cuh"
#include "sshex8_laplacian_core.cuh"
#include "sshex8_laplacian_stream.cuh"
#include "sshex8_laplacian_tma.cuh"

namespace b200fem {

// Slot-plane buffer layout (elements): the interior run of plane z sits at offset k = g & (AL - 1), g = the node id
// it starts with, AL = 16 / sizeof(T): the innermost tensor coordinate of a TMA copy must be a 16-byte multiple
// (measured: tools/dev/tma1d_probe.cu -- odd fp64 coordinates raise "illegal instruction"), so the box starts at
// g - k. Consecutive elements of a compact numbering have interior blocks of (L-1)^3 = 343 ids, i.e. k alternates
// between the slots of a group -- which is exactly the skew that keeps the x-rows of neighbouring slots on disjoint
// banks. The skeleton block of slot s starts at SK(s) + DK(s) k: with a fixed start one of the two parity patterns of
// a group, (k0, k1, k2) = (1, 0, 1), costs 89 wavefronts for the 27 loads of a z-step and 35 for the 9 stores (ideal
// 54 / 18; ncu: 2.8 wavefronts per LDS.64); the k-dependent start brings both patterns to 60 / 20 (exhaustive search,
// same model).
template <typename T>
struct WsLayout;
template <>
struct WsLayout<double> {
    static constexpr int BOX = 50;  // elements per TMA box (400 B) >= (L-1)^2 + AL - 1
    // skeleton block of slot s at SK(s) + DK(s) * k, k = the alignment offset of that slot's interior run in this plane
    __host__ __device__ static constexpr int SK(int s) { return s == 0 ? 56 : (s == 1 ? 57 : 53); }
    __host__ __device__ static constexpr int DK(int s) { return s == 0 ? 0 : (s == 1 ? 1 : -1); }
};
template <>
struct WsLayout<float> {
    static constexpr int BOX = 52;  // 208 B
    __host__ __device__ static constexpr int SK(int s) { return s == 0 ? 52 : (s == 1 ? 56 : 59); }
    __host__ __device__ static constexpr int DK(int) { return 0; }
};

template <typename T, int L, int NC, int D, int S, int FACEOUT = 0>
struct Sshex8WsCfg {
    static constexpr int N = L + 1;
    static constexpr int NI = (L - 1) * (L - 1);  // interior positions of a plane
    static constexpr int NSK = N * N - NI;        // skeleton positions of a plane (4 L)
    static constexpr int EPT = 32 / N;            // elements per compute warp at a time
    static constexpr int PS = 96;                 // elements per slot-plane buffer (128-byte multiple for 4 and 8 B)
    static constexpr int PLANE = EPT * PS;
    static constexpr size_t ring_bytes = (size_t)D * PLANE * sizeof(T);
    static constexpr size_t stage_bytes = (size_t)S * PLANE * sizeof(T);
    static constexpr size_t hdr_bytes = (size_t)3 * EPT * sizeof(Sshex8Header);
    static constexpr size_t bar_bytes = (size_t)(2 * D + 2 * S) * sizeof(uint64_t);
    // elements per slot of the lateral output buffer. FACEOUT 1: all 32 skeleton positions x (L-1) planes = 224 used (228: 3
    // wavefronts per store, bank model); FACEOUT 2: the 18 positions with x = 0 or x = L only = 126 used (129)
    static constexpr int FOS = FACEOUT == 2 ? 129 : 228;
    static constexpr size_t fo_bytes = FACEOUT ? (size_t)EPT * FOS * sizeof(T) : 0;
    static constexpr size_t warp_bytes = ((ring_bytes + stage_bytes + hdr_bytes + bar_bytes + fo_bytes + 127) / 128) * 128;
    static constexpr size_t smem_bytes = warp_bytes * NC;
    static constexpr int threads = 2 * NC * 32;
    static_assert(L == 8, "the ws kernel is built for level 8 (EPT = 3, 32 skeleton positions per plane)");
    static_assert(NSK == 32, "one skeleton position per lane and slot");
    static_assert(WsLayout<T>::BOX * sizeof(T) % 16 == 0, "TMA box must be a multiple of 16 bytes");
};

__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
// arrive on `bar` once all cp.async copies this thread has issued so far have landed (counts against the init count)
__device__ __forceinline__ void cp_async_mbar_arrive_noinc(uint64_t *bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void *smem_dst, const CUtensorMap *tmap, int c0, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.1d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2}], [%3];\n" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(reinterpret_cast<uint64_t>(tmap)), "r"(c0), "r"(smem_u32(bar))
                 : "memory");
}

// offset of a skeleton position of slot i relative to the same position of slot 0
template <typename T>
__device__ __forceinline__ constexpr int ws_slot_shift(const int i) {
    return i * 96 + WsLayout<T>::SK(i) - WsLayout<T>::SK(0);
}

// element offset of lattice position (x, yy) inside the buffer of slot s (interior positions: before the per-plane
// alignment offset k is added)
template <typename T, int L>
__device__ __forceinline__ int ws_offset(const int s, const int x, const int yy) {
    constexpr int PS = 96;
    constexpr int k = 0;
    const int sk = (s == 0) ? WsLayout<T>::SK(0) : (s == 1 ? WsLayout<T>::SK(1) : WsLayout<T>::SK(2));
    // (yy may come in unclamped: rows below 0 / above L are the boundary rows. Written with <= / >= on purpose:
    // ptxas 12.9 folded a clamp followed by `== L` into the predicate output of VIMNMX and took the `yy == L`
    // branch for every lane.)
    int o;
    if (yy <= 0)
        o = sk + x;
    else if (yy >= L)
        o = sk + (L + 1) + x;
    else if (x == 0)
        o = sk + 2 * (L + 1) + (yy - 1);
    else if (x == L)
        o = sk + 2 * (L + 1) + (L - 1) + (yy - 1);
    else
        o = k + (yy - 1) * (L - 1) + (x - 1);
    return s * PS + o;
}

template <typename T, int L, int NC, int D, int S, bool PLAIN_INTERIOR, bool PHASED>
__global__ void __launch_bounds__(2 * NC * 32, 1)
    sshex8_laplacian_ws_kernel(const ptrdiff_t ne, const Sshex8Header *__restrict__ headers, const __grid_constant__ CUtensorMap tmap_u,
                               const T *__restrict__ u, T *__restrict__ y, const int dbg) {
    using Cfg = Sshex8WsCfg<T, L, NC, D, S>;
    using Lay = WsLayout<T>;
    constexpr int N = Cfg::N, NI = Cfg::NI, EPT = Cfg::EPT, PS = Cfg::PS, PLANE = Cfg::PLANE;
    constexpr int HSZ = (int)sizeof(Sshex8Header);
    constexpr int HBUF = EPT * HSZ;
    constexpr uint32_t TX_BYTES = (uint32_t)(EPT * Lay::BOX * sizeof(T));
    constexpr int AL = 16 / (int)sizeof(T);  // elements per 16 bytes
    // register budgets: 2 NC warps start with 65536 / (64 NC) each
    constexpr int REG_MOVER = 80;
    constexpr int REG_COMPUTE = (65536 / 32 - NC * REG_MOVER) / NC / 8 * 8;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, tl = threadIdx.x & 31;
    const bool is_mover = warp >= NC;
    const int cw = is_mover ? warp - NC : warp;  // the compute warp this warp is, or serves
    unsigned char *wbase = smem_raw + (size_t)cw * Cfg::warp_bytes;
    T *ring = reinterpret_cast<T *>(wbase);                                      // [D][EPT][PS]
    T *stage = ring + (size_t)D * PLANE;                                         // [S][EPT][PS]
    unsigned char *hdr_bytes = wbase + Cfg::ring_bytes + Cfg::stage_bytes;       // [3][EPT] headers
    uint64_t *bars = reinterpret_cast<uint64_t *>(hdr_bytes + Cfg::hdr_bytes);
    uint64_t *full = bars, *empty = bars + D, *sfull = bars + 2 * D, *sempty = bars + 2 * D + S;

    // ---- work of this warp pair: groups gteam, gteam + nteams, ... The last group may hold fewer than EPT elements: its
    // missing slots read the padding headers (copies of the last element: same loads, same arithmetic) and are not
    // scattered.
    const int ngroups = (int)((ne + EPT - 1) / EPT);  // element ids are int32 (idx_t)
    const int nteams = (int)gridDim.x * NC;
    const int gteam = (int)blockIdx.x * NC + cw;
    const int my_groups = gteam < ngroups ? (ngroups - gteam + nteams - 1) / nteams : 0;
    const int T_steps = my_groups * N;

    if (!is_mover && tl == 0) {
#pragma unroll
        for (int i = 0; i < D; ++i) {
            mbar_init(full + i, 33);  // 32 cp.async-triggered arrivals + the elected lane's (expect_tx) arrival
            mbar_init(empty + i, 1);
        }
#pragma unroll
        for (int i = 0; i < S; ++i) {
            mbar_init(sfull + i, 1);
            mbar_init(sempty + i, 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    fence_proxy_async_smem();
    __syncthreads();

    auto group_of = [&](int g_local) -> int { return gteam + g_local * nteams; };

    if (is_mover) {
        // =========================================================================================== mover warp
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;\n" ::"n"(REG_MOVER));
        if (my_groups == 0) return;

        // skeleton item i: slot i, plane position NI + tl (interior-first order of plane_pos_to_xy)
        int sk_x, sk_y;
        plane_pos_to_xy<L>(NI + tl, sk_x, sk_y);
        const int sk_cxy = ((sk_x == 0) ? 0 : (sk_x == L ? 2 : 1)) + 3 * ((sk_y == 0) ? 0 : (sk_y == L ? 2 : 1));
        const int sk_off0 = ws_offset<T, L>(0, sk_x, sk_y);  // slot i: + ws_slot_shift<T>(i)
        // interior items of slot s: run positions q = tl (always) and q = 32 + tl (lanes below NI - 32); the slot is a
        // compile-time index, so per plane and slot there is one base id / base offset and the items are one add away
        const bool in_second = tl < NI - 32;
        const int q1 = in_second ? 32 + tl : 0;
        const int x0 = 1 + tl % (L - 1), y0 = 1 + tl / (L - 1), x1 = 1 + q1 % (L - 1), y1 = 1 + q1 / (L - 1);

        auto issue_headers = [&](int g_local) {
            if (g_local >= my_groups) return;
            const int4 *src = reinterpret_cast<const int4 *>(headers + (size_t)group_of(g_local) * EPT);
            int4 *dst = reinterpret_cast<int4 *>(hdr_bytes + (g_local % 3) * HBUF);
            for (int i = tl; i < EPT * 29; i += 32) cp_async_16(dst + i, src + i);
        };
        // skeleton ids of the z-interior planes of a group: id(z) = I + S z; interior run of plane z starts at B + NI z
        auto plane_set = [&](const unsigned char *hb, int (&I)[EPT], int (&Sz)[EPT], int (&B)[EPT]) {
#pragma unroll
            for (int i = 0; i < EPT; ++i) {
                const int4 d = *reinterpret_cast<const int4 *>(hb + i * HSZ + (9 + sk_cxy) * 16);
                I[i] = d.x + d.y * sk_x + d.z * sk_y;
                Sz[i] = d.w;
                const int4 c = *reinterpret_cast<const int4 *>(hb + i * HSZ + 13 * 16);
                B[i] = c.x + c.y + c.z;  // id(1, 1, z) = B + sz z with the canonical strides (1, L-1, (L-1)^2)
            }
        };
        // node id on a boundary plane (cz = 0 / 2): the position's own box descriptor
        auto boundary_id = [&](const unsigned char *hb, int slot, int x, int yy, int cxy, int z) -> int {
            const int4 d = *reinterpret_cast<const int4 *>(hb + slot * HSZ + ((z == 0 ? 0 : 18) + cxy) * 16);
            return d.x + d.y * x + d.z * yy + d.w * z;
        };
        auto put = [&](T *p, const T v, auto plain_tag) {  // streaming store or RED, chosen at compile time
            if constexpr (decltype(plain_tag)::value) {
                if constexpr (sizeof(T) == 8)
                    asm volatile("st.global.cs.f64 [%0], %1;\n" ::"l"(p), "d"(v));
                else
                    asm volatile("st.global.cs.f32 [%0], %1;\n" ::"l"(p), "f"(v));
            } else {
                if constexpr (sizeof(T) == 8)
                    asm volatile("red.global.add.f64 [%0], %1;\n" ::"l"(p), "d"(v));
                else
                    asm volatile("red.global.add.f32 [%0], %1;\n" ::"l"(p), "f"(v));
            }
        };
        using PlainT = std::integral_constant<bool, PLAIN_INTERIOR>;
        using RedT = std::integral_constant<bool, false>;

        int ldI[EPT], ldS[EPT], ldB[EPT];  // set of the group whose planes are being gathered
        int scI[EPT], scS[EPT], scB[EPT];  // set of the group whose planes are being scattered

        // gather plane zp of the group with header buffer hb into ring slot `rs` (the slot is free)
        auto issue_plane = [&](const int rs, const unsigned char *hb, const int zp) {
            uint64_t *bar = full + rs;
            T *dst = ring + (size_t)rs * PLANE;
            if (zp == 0) {
                // headers of this group were issued N commit groups ago
                cp_async_wait<N - 2>();
                __syncwarp();
                plane_set(hb, ldI, ldS, ldB);
            }
            if (zp == 0 || zp == L) {
                if (tl == 0) mbar_arrive(bar);
#pragma unroll
                for (int i = 0; i < EPT; ++i) {
                    const int k = (ldB[i] + NI * zp) & (AL - 1);
                    cp_async_value(dst + (sk_off0 + ws_slot_shift<T>(i) + Lay::DK(i) * k), u + boundary_id(hb, i, sk_x, sk_y, sk_cxy, zp));
                    cp_async_value(dst + (i * PS + k + tl), u + boundary_id(hb, i, x0, y0, 4, zp));
                    if (in_second) cp_async_value(dst + (i * PS + k + q1), u + boundary_id(hb, i, x1, y1, 4, zp));
                    asm volatile("" ::: "memory");  // one slot at a time: small register budget
                }
            } else if (dbg & 1) {  // A/B switch: interior runs per node as well (no TMA)
                if (tl == 0) mbar_arrive(bar);
#pragma unroll
                for (int i = 0; i < EPT; ++i) {
                    const int g = ldB[i] + NI * zp, k = g & (AL - 1);
                    cp_async_value(dst + (sk_off0 + ws_slot_shift<T>(i) + Lay::DK(i) * k), u + (ldI[i] + ldS[i] * zp));
                    cp_async_value(dst + (i * PS + k + tl), u + (g + tl));
                    if (in_second) cp_async_value(dst + (i * PS + k + q1), u + (g + q1));
                }
            } else {
                if (tl == 0) {
                    mbar_arrive_expect_tx(bar, TX_BYTES);
#pragma unroll
                    for (int i = 0; i < EPT; ++i) {
                        const int g = ldB[i] + NI * zp;
                        tma_load_1d(dst + i * PS, &tmap_u, g - (g & (AL - 1)), bar);
                    }
                }
#pragma unroll
                for (int i = 0; i < EPT; ++i) {
                    const int k = (ldB[i] + NI * zp) & (AL - 1);
                    cp_async_value(dst + (sk_off0 + ws_slot_shift<T>(i) + Lay::DK(i) * k), u + (ldI[i] + ldS[i] * zp));
                }
            }
            cp_async_mbar_arrive_noinc(bar);
        };

        // output plane z of the group with header buffer hb, staged in slot ss (phase parity sp); the staging slot is
        // handed back after the last value has been read
        auto scatter_plane = [&](const int ss, const uint32_t sp, const unsigned char *hb, const int z, const int nv) {
            const T *src = stage + (size_t)ss * PLANE;
            mbar_wait(sfull + ss, sp);
            if (z == 0 || z == L) {
#pragma unroll
                for (int i = 0; i < EPT; ++i) {
                    if (i >= nv) break;
                    const int k = (scB[i] + NI * z) & (AL - 1);
                    put(y + boundary_id(hb, i, sk_x, sk_y, sk_cxy, z), src[sk_off0 + ws_slot_shift<T>(i) + Lay::DK(i) * k], RedT{});
                    put(y + boundary_id(hb, i, x0, y0, 4, z), src[i * PS + k + tl], RedT{});
                    if (in_second) put(y + boundary_id(hb, i, x1, y1, 4, z), src[i * PS + k + q1], RedT{});
                    asm volatile("" ::: "memory");
                }
            } else {
#pragma unroll
                for (int i = 0; i < EPT; ++i) {
                    if (i >= nv) break;
                    const int g = scB[i] + NI * z, k = g & (AL - 1);
                    put(y + (scI[i] + scS[i] * z), src[sk_off0 + ws_slot_shift<T>(i) + Lay::DK(i) * k], RedT{});
                    put(y + (g + tl), src[i * PS + k + tl], PlainT{});
                    if (in_second) put(y + (g + q1), src[i * PS + k + q1], PlainT{});
                }
            }
            __syncwarp();
            if (tl == 0) mbar_arrive(sempty + ss);
        };

        // ---- prologue: headers of groups 0 and 1, planes 0 .. D-1
        issue_headers(0);
        cp_async_commit();
        cp_async_wait<0>();
        __syncwarp();
        issue_headers(1);
#pragma unroll 1
        for (int s = 0; s < D; ++s) {
            if (s < T_steps) issue_plane(s, hdr_bytes, s);  // D <= N: all in group 0
            cp_async_commit();
        }
        // ---- stream. Counters instead of divisions: (gn, zn, hn) = group, plane and header buffer of step t + D;
        // (zp, hs) = plane and header buffer of step t; rs / re = ring slot and `empty` parity of step t + D;
        // ss / sp = staging slot and `sfull` parity of the next output plane.
        int gn = 0, zn = D, hn = 0, rs = 0;
        uint32_t re = 0;
        if (zn == N) { zn = 0; gn = 1; hn = 1; }
        int zp = 0, hs = 0, ss = 0, gs = 0, nv = EPT;
        uint32_t sp = 0;
        static_assert(D <= N, "ring deeper than a group");
#pragma unroll 1
        for (int t = 0; t < T_steps; ++t) {
            if (t + D < T_steps) {
                mbar_wait(empty + rs, re);
                if (zn == 0) issue_headers(gn + 1);  // headers of group g+2 travel with the first plane of g+1
                issue_plane(rs, hdr_bytes + hn * HBUF, zn);
            }
            cp_async_commit();
            if (++rs == D) { rs = 0; re ^= 1u; }
            if (++zn == N) { zn = 0; ++gn; hn = (hn == 2) ? 0 : hn + 1; }

            const unsigned char *hb = hdr_bytes + hs * HBUF;
            if (zp == 0) {
                plane_set(hb, scI, scS, scB);
                const int left = (int)ne - group_of(gs) * EPT;  // valid slots of the group being scattered
                nv = left < EPT ? left : EPT;
            }
            if (zp > 0) {
                scatter_plane(ss, sp, hb, zp - 1, nv);
                if (++ss == S) { ss = 0; sp ^= 1u; }
            }
            if (zp == L) {
                scatter_plane(ss, sp, hb, L, nv);
                if (++ss == S) { ss = 0; sp ^= 1u; }
            }
            if (++zp == N) { zp = 0; ++gs; hs = (hs == 2) ? 0 : hs + 1; }
        }
        cp_async_wait<0>();
    } else {
        // ========================================================================================= compute warp
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;\n" ::"n"(REG_COMPUTE));
        if (my_groups == 0) return;
        // lanes beyond EPT * N shadow the last row (same addresses, same values: no divergent region)
        const int cl = tl < EPT * N ? tl : EPT * N - 1;
        const int slot = cl / N;
        const int yrow = cl - slot * N;
        // the three rows this lane reads: offsets of x = 0 and x = 1 (x = 1 .. L-1 are contiguous; x = L follows from x = 0)
        int a0[3], am[3];
#pragma unroll
        for (int iy = 0; iy < 3; ++iy) {
            const int yy = yrow + iy - 1;  // unclamped, see ws_offset
            a0[iy] = ws_offset<T, L>(slot, 0, yy);
            am[iy] = ws_offset<T, L>(slot, 1, yy);
        }
        // rows of the interior block carry the per-plane alignment offset of the run (see WsLayout)
        bool rint[3];
#pragma unroll
        for (int iy = 0; iy < 3; ++iy) {
            const int yy = yrow + iy - 1;
            rint[iy] = (yy >= 1 && yy <= L - 1);
        }
        const int dk = (slot == 0) ? Lay::DK(0) : (slot == 1 ? Lay::DK(1) : Lay::DK(2));  // skeleton block moves by dk * k
        int Bs = 0;  // id of lattice node (1, 1, 0) of this lane's element: the interior run of plane z starts at Bs + NI z
        Sshex8LaplacianRow<T, L> row;

        auto emit = [&](int o, int z, const T(&vals)[N]) {  // stage output plane z (number o overall), this lane's row
            if (o >= S) mbar_wait(sempty + (o % S), (uint32_t)(((o / S) - 1) & 1));
            T *dst = stage + (size_t)(o % S) * PLANE;
            const int k = (Bs + NI * z) & (AL - 1);
            T *ds = dst + dk * k;                        // skeleton positions
            T *dm = (rint[1] ? dst + k : ds) + am[1];    // x = 1 .. L-1 of this lane's row
            ds[a0[1]] = vals[0];
#pragma unroll
            for (int x = 1; x < L; ++x) dm[x - 1] = vals[x];
            ds[a0[1] + (rint[1] ? L - 1 : L)] = vals[L];
            __syncwarp();
            if (tl == 0) mbar_arrive(sfull + (o % S));
        };

        // One z-step with compile-time plane class ZC (0: z = 0, 1: interior, 2: z = L) and accumulator phase P = zp % 3
        // (plane z accumulates in row.acc[z % 3]: no register rotation; N = 9 is a multiple of 3, so the phase pattern
        // is the same for every group).
        auto z_step = [&](auto zc_tag, auto p_tag, const int g_local, const int zp) {
            constexpr int ZC = decltype(zc_tag)::value, P = decltype(p_tag)::value;
            const int t = g_local * N + zp;
            const T *plane = ring + (size_t)(t % D) * PLANE;
            mbar_wait(full + (t % D), (uint32_t)((t / D) & 1));
            if constexpr (ZC == 0) {
                const Sshex8Header *h = reinterpret_cast<const Sshex8Header *>(hdr_bytes + (g_local % 3) * HBUF) + slot;
                T f[6];
#pragma unroll
                for (int k = 0; k < 6; ++k) f[k] = (T)h->fff[k];
                row.init(f, yrow);
                Bs = h->box[13][0] + h->box[13][1] + h->box[13][2];
            }
            const int kin = (Bs + NI * zp) & (AL - 1);
            const T *pk = plane + kin, *ps = plane + dk * kin;  // interior rows / skeleton positions
            row.template step_class_with<ZC, P>([&](const int iy, T(&r)[N]) {
                const T *pm = (rint[iy] ? pk : ps) + am[iy];
                r[0] = ps[a0[iy]];
#pragma unroll
                for (int x = 1; x < L; ++x) r[x] = pm[x - 1];
                r[L] = ps[a0[iy] + (rint[iy] ? L - 1 : L)];  // x = L sits L-1 (column block) or L (boundary row) after x = 0
            });
            __syncwarp();
            if (tl == 0) mbar_arrive(empty + (t % D));
            if constexpr (ZC != 0) {  // output plane zp - 1 is final; its registers become the accumulator of plane zp + 2
                emit(t - 1, zp - 1, row.acc[(P + 2) % 3]);
#pragma unroll
                for (int x = 0; x < N; ++x) row.acc[(P + 2) % 3][x] = T(0);
            }
            if constexpr (ZC == 2) emit(t, L, row.acc[P % 3]);
        };
        using I0 = std::integral_constant<int, 0>;
        using I1 = std::integral_constant<int, 1>;
        using I2 = std::integral_constant<int, 2>;
        static_assert(N % 3 == 0, "phase pattern assumes N % 3 == 0");
        if constexpr (!PHASED) {
            // rolled form: one loop body, accumulators rotated through the registers after every step
            int g_local = 0, zp = 0;
#pragma unroll 1
            for (int t = 0; t < T_steps; ++t) {
                const T *plane = ring + (size_t)(t % D) * PLANE;
                mbar_wait(full + (t % D), (uint32_t)((t / D) & 1));
                if (zp == 0) {
                    const Sshex8Header *h = reinterpret_cast<const Sshex8Header *>(hdr_bytes + (g_local % 3) * HBUF) + slot;
                    T f[6];
#pragma unroll
                    for (int k = 0; k < 6; ++k) f[k] = (T)h->fff[k];
                    row.init(f, yrow);
                    Bs = h->box[13][0] + h->box[13][1] + h->box[13][2];
                }
                const int kin = (Bs + NI * zp) & (AL - 1);
                const T *pk = plane + kin, *ps = plane + dk * kin;  // interior rows / skeleton positions
                row.step_with(zp, [&](const int iy, T(&r)[N]) {
                    const T *pm = (rint[iy] ? pk : ps) + am[iy];
                    r[0] = ps[a0[iy]];
#pragma unroll
                    for (int x = 1; x < L; ++x) r[x] = pm[x - 1];
                    r[L] = ps[a0[iy] + (rint[iy] ? L - 1 : L)];  // x = L sits L-1 (column block) or L (boundary row) after x = 0
                });
                __syncwarp();
                if (tl == 0) mbar_arrive(empty + (t % D));
                if (zp > 0) emit(g_local * N + zp - 1, zp - 1, row.acc[0]);
                row.rotate();
                if (zp == L) emit(g_local * N + L, L, row.acc[0]);
                if (++zp == N) {
                    zp = 0;
                    ++g_local;
                }
            }
            return;
        }
#pragma unroll 1
        for (int g_local = 0; g_local < my_groups; ++g_local) {
            z_step(I0{}, I0{}, g_local, 0);
#pragma unroll 1
            for (int zz = 0; zz < N; zz += 3) {
                if (zz > 0) z_step(I1{}, I0{}, g_local, zz);
                z_step(I1{}, I1{}, g_local, zz + 1);
                if (zz + 2 < L)
                    z_step(I1{}, I2{}, g_local, zz + 2);
                else
                    z_step(I2{}, I2{}, g_local, L);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// "ws2": the same roles, layout and arithmetic with HALF the hand-offs. Bracketing the first kernel with its parts switched
// off (profiles/r02p_ws_brackets.jsonl, C3 fp64): everything 0.86 ms; arithmetic alone (movers only run the barrier protocol)
// 0.73 ms; loads + stores alone 0.68 ms; the barrier protocol ALONE 0.29 ms. A DFMA stream needs 2 cycles per instruction and
// scheduler (tools/dev/dfma_probe.cu: one warp with >= 4 independent chains saturates the pipe), i.e. 0.30 ms for this kernel:
// the four waits / arrivals a warp pair exchanges per plane cost as much as the arithmetic. Here
//   * the compute warp emits output plane t-1 after EVERY step t (the last plane of a group rides into the first step of the
//     next group: ZC = 0 steps do not touch that accumulator), so `sfull` of step t also says "ring slot t % D has been
//     read": no `empty` barrier;
//   * the mover issues the loads of plane t + D, scatters plane t - 1, and only then adds the last arrival to full[t % D]:
//     with as many staging slots as ring slots, `full` of plane t says "staging slot t % D has been scattered": no `sempty`
//     barrier either (a phantom plane T covers the final emit);
//   * all 32 lanes arrive on `sfull` (no __syncwarp + elected arrive).
// Tried and dropped: the interior output runs by cp.async.bulk (48 of the 49 values of a run are 16-byte aligned) instead of two
// coalesced stores per slot -- 0.99 against 0.85 ms: the mover has to wait for the bulk engine to have read the staging slot
// before it may hand the slot back (profiles/r02q_ws2.jsonl and the runs around it).
// try_wait with a suspend-time hint: the warp sleeps in hardware until the phase completes (or the hint runs out) instead of
// re-issuing the poll -- the polls of a waiting role take issue slots from the role it waits for
template <int HINT_NS>
__device__ __forceinline__ void mbar_wait_hint(uint64_t *bar, uint32_t parity) {
    if constexpr (HINT_NS == 0) {
        mbar_wait(bar, parity);
    } else {
        asm volatile(
            "{\n .reg .pred p;\n WAITH_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n @p bra DONEH_%=;\n bra WAITH_%=;\n DONEH_%=:\n}\n" ::"r"(
                smem_u32(bar)),
            "r"(parity), "r"(HINT_NS)
            : "memory");
    }
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

template <typename T, int L, int NC, int D, bool PLAIN_INTERIOR, bool PHASED, int DBG = 0, int FACEOUT = 0, int HINT = 0>
__global__ void __launch_bounds__(2 * NC * 32, 1)
    sshex8_laplacian_ws2_kernel(const ptrdiff_t ne, const Sshex8Header *__restrict__ headers, const __grid_constant__ CUtensorMap tmap_u,
                                const T *__restrict__ u, T *__restrict__ y, const WsDependency dep) {
    constexpr int dbg = DBG;  // experiment switches (bracketing runs): 2 = no arithmetic, 4 = no stores, 8 = no loads, 16 = no REDs, 32 = no plain stores
    using Cfg = Sshex8WsCfg<T, L, NC, D, D, FACEOUT>;
    using Lay = WsLayout<T>;
    constexpr int FOS = Cfg::FOS;
    constexpr int N = Cfg::N, NI = Cfg::NI, EPT = Cfg::EPT, PS = Cfg::PS, PLANE = Cfg::PLANE;
    constexpr int HSZ = (int)sizeof(Sshex8Header);
    constexpr int HBUF = EPT * HSZ;
    constexpr uint32_t TX_BYTES = (uint32_t)(EPT * Lay::BOX * sizeof(T));
    constexpr int AL = 16 / (int)sizeof(T);
    constexpr int REG_MOVER = 80;
    constexpr int REG_COMPUTE = (65536 / 32 - NC * REG_MOVER) / NC / 8 * 8;
    // the register pool of the CTA is threads x the launch allocation, which is 65536 only for NC = 8 (128 registers at launch):
    // with more warp pairs setmaxnreg.inc would wait for registers that never come
    static_assert(NC == 8, "register budgets are derived for 8 warp pairs (512 threads x 128 registers)");
    static_assert(D < N, "ring deeper than a group");

    extern __shared__ __align__(128) unsigned char smem_raw[];
    // (the shuffle tells the compiler that the warp index, and everything derived from it, is warp-uniform)
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), tl = threadIdx.x & 31;
    const bool is_mover = warp >= NC;
    const int cw = is_mover ? warp - NC : warp;
    unsigned char *wbase = smem_raw + (size_t)cw * Cfg::warp_bytes;
    T *ring = reinterpret_cast<T *>(wbase);                                 // [D][EPT][PS]
    T *stage = ring + (size_t)D * PLANE;                                    // [D][EPT][PS]
    unsigned char *hdr_bytes = wbase + Cfg::ring_bytes + Cfg::stage_bytes;  // [3][EPT] headers
    uint64_t *bars = reinterpret_cast<uint64_t *>(hdr_bytes + Cfg::hdr_bytes);
    uint64_t *full = bars, *sfull = bars + D;
    // FACEOUT: the skeleton values of the z-interior output planes of the current group, [slot][32 positions][L-1 planes]:
    // written by the compute warp plane after plane, reduced into y by the mover in two sweeps per group with z fastest --
    // the lateral faces of the reference numbering are contiguous along z, so a sweep instruction touches 8 - 11 sectors where
    // the 32 positions of one plane are 32 sectors (profiles/r02*_ws2_brackets: the store path, 101 M single-sector REDs, is
    // what bounds the kernel once the hand-offs are cheap)
    T *fo = reinterpret_cast<T *>(wbase + Cfg::ring_bytes + Cfg::stage_bytes + Cfg::hdr_bytes + Cfg::bar_bytes);

    const int ngroups = (int)((ne + EPT - 1) / EPT);
    const int nteams = (int)gridDim.x * NC;
    const int gteam = (int)blockIdx.x * NC + cw;
    const int my_groups = gteam < ngroups ? (ngroups - gteam + nteams - 1) / nteams : 0;
    const int T_steps = my_groups * N;

    if (!is_mover && tl == 0) {
#pragma unroll
        for (int i = 0; i < D; ++i) {
            mbar_init(full + i, 33);   // 32 cp.async-triggered arrivals + the mover's closing arrival
            mbar_init(sfull + i, 32);  // every lane of the compute warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    fence_proxy_async_smem();
    __syncthreads();

    auto group_of = [&](int g_local) -> int { return gteam + g_local * nteams; };

    if (is_mover) {
        // =========================================================================================== mover warp
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;\n" ::"n"(REG_MOVER));
        if (my_groups == 0) return;

        int sk_x, sk_y;
        plane_pos_to_xy<L>(NI + tl, sk_x, sk_y);
        const int sk_cxy = ((sk_x == 0) ? 0 : (sk_x == L ? 2 : 1)) + 3 * ((sk_y == 0) ? 0 : (sk_y == L ? 2 : 1));
        const int sk_off0 = ws_offset<T, L>(0, sk_x, sk_y);
        const bool in_second = tl < NI - 32;
        const int q1 = in_second ? 32 + tl : 0;
        const int x0 = 1 + tl % (L - 1), y0 = 1 + tl / (L - 1), x1 = 1 + q1 % (L - 1), y1 = 1 + q1 / (L - 1);

        auto issue_headers = [&](int g_local) {
            if (g_local >= my_groups) return;
            const int4 *src = reinterpret_cast<const int4 *>(headers + (size_t)group_of(g_local) * EPT);
            int4 *dst = reinterpret_cast<int4 *>(hdr_bytes + (g_local % 3) * HBUF);
            for (int i = tl; i < EPT * 29; i += 32) cp_async_16(dst + i, src + i);
        };
        // per group and slot: G = id of the first interior node of plane 0 (the interior run of plane z starts at G + NI z: kept
        // as a running value), I / Sz = id and z stride of this lane's skeleton position on the z-interior planes
        auto plane_set = [&](const unsigned char *hb, int (&I)[EPT], int (&Sz)[EPT], int (&G)[EPT]) {
#pragma unroll
            for (int i = 0; i < EPT; ++i) {
                const int4 d = *reinterpret_cast<const int4 *>(hb + i * HSZ + (9 + sk_cxy) * 16);
                I[i] = d.x + d.y * sk_x + d.z * sk_y;
                Sz[i] = d.w;
                const int4 c = *reinterpret_cast<const int4 *>(hb + i * HSZ + 13 * 16);
                G[i] = c.x + c.y + c.z;  // id(1, 1, z) = G + NI z with the canonical strides (1, L-1, (L-1)^2)
            }
        };
        auto boundary_id = [&](const unsigned char *hb, int slot, int x, int yy, int cxy, int z) -> int {
            const int4 d = *reinterpret_cast<const int4 *>(hb + slot * HSZ + ((z == 0 ? 0 : 18) + cxy) * 16);
            return d.x + d.y * x + d.z * yy + d.w * z;
        };
        // predicated forms (no branch, no reconvergence point around a store / copy)
        auto put_if = [&](const bool ok, T *p, const T v, auto plain_tag) {
            if (dbg & (decltype(plain_tag)::value ? 32 : 16)) return;  // experiment: no plain stores / no REDs
            if constexpr (decltype(plain_tag)::value) {
                if constexpr (sizeof(T) == 8)
                    asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %2, 0;\n @p st.global.cs.f64 [%0], %1;\n}\n" ::"l"(p), "d"(v), "r"((int)ok));
                else
                    asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %2, 0;\n @p st.global.cs.f32 [%0], %1;\n}\n" ::"l"(p), "f"(v), "r"((int)ok));
            } else {
                if constexpr (sizeof(T) == 8)
                    asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %2, 0;\n @p red.global.add.f64 [%0], %1;\n}\n" ::"l"(p), "d"(v), "r"((int)ok));
                else
                    asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %2, 0;\n @p red.global.add.f32 [%0], %1;\n}\n" ::"l"(p), "f"(v), "r"((int)ok));
            }
        };
        auto copy_if = [&](const bool ok, T *smem, const T *gmem) {
            if constexpr (sizeof(T) == 8)
                asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %2, 0;\n @p cp.async.ca.shared.global [%0], [%1], 8;\n}\n" ::"r"(smem_u32(smem)), "l"(gmem),
                             "r"((int)ok)
                             : "memory");
            else
                asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %2, 0;\n @p cp.async.ca.shared.global [%0], [%1], 4;\n}\n" ::"r"(smem_u32(smem)), "l"(gmem),
                             "r"((int)ok)
                             : "memory");
        };
        using PlainT = std::integral_constant<bool, PLAIN_INTERIOR>;
        using RedT = std::integral_constant<bool, false>;

        // sweep items of this lane: half A = planes 1 .. 4 (4 items: position (tl >> 2) + 8 q, plane 1 + (tl & 3)), half B = planes
        // 5 .. L-1 (3 items: position j / 3, plane 5 + j % 3, j = tl + 32 q), packed as box byte offset | x << 9 | y << 13 |
        // buffer offset << 17 | plane << 25
        int sweepA[4], sweepB[3];
        constexpr int NSA = FACEOUT == 2 ? 3 : 4, NSB = FACEOUT == 2 ? 2 : 3;  // RED instructions per slot and half
        if constexpr (FACEOUT == 1) {
            auto pack = [&](int pos, int zpl) {
                int x, yy;
                plane_pos_to_xy<L>(NI + pos, x, yy);
                const int cxy = ((x == 0) ? 0 : (x == L ? 2 : 1)) + 3 * ((yy == 0) ? 0 : (yy == L ? 2 : 1));
                return ((9 + cxy) * 16) | (x << 9) | (yy << 13) | ((pos * (L - 1) + zpl - 1) << 17) | (zpl << 25) | (1 << 28);
            };
#pragma unroll
            for (int q = 0; q < 4; ++q) sweepA[q] = pack((tl >> 2) + 8 * q, 1 + (tl & 3));
#pragma unroll
            for (int q = 0; q < 3; ++q) sweepB[q] = pack((tl + 32 * q) / 3, 5 + (tl + 32 * q) % 3);
        }
        if constexpr (FACEOUT == 2) {
            // FACEOUT 2: only the 18 positions with x = 0 or x = L (the columns of the lateral x-faces + the four z-edges) go
            // through the buffer -- a compute lane writes them as two scalars, at no cost in its staging code; the rows y = 0 / L
            // keep the per-plane path. Buffer position c: 0..3 = (0,0) (L,0) (0,L) (L,L), 4..10 = (0, 1..L-1), 11..17 = (L, 1..L-1)
            auto pack2 = [&](int j, int nval, int c, int zpl) {
                c = c < 18 ? c : 17;
                const int x = (c < 4) ? ((c & 1) ? L : 0) : (c < 11 ? 0 : L);
                const int yy = (c < 4) ? ((c & 2) ? L : 0) : (c < 11 ? c - 3 : c - 10);
                const int cxy = ((x == 0) ? 0 : 2) + 3 * ((yy == 0) ? 0 : (yy == L ? 2 : 1));
                return ((9 + cxy) * 16) | (x << 9) | (yy << 13) | ((c * (L - 1) + zpl - 1) << 17) | (zpl << 25) | ((j < nval ? 1 : 0) << 28);
            };
#pragma unroll
            for (int q = 0; q < 3; ++q) sweepA[q] = pack2(tl + 32 * q, 18 * 4, (tl + 32 * q) >> 2, 1 + ((tl + 32 * q) & 3));
#pragma unroll
            for (int q = 0; q < 2; ++q) sweepB[q] = pack2(tl + 32 * q, 18 * 3, (tl + 32 * q) / 3, 5 + (tl + 32 * q) % 3);
        }
        auto sweep = [&](const unsigned char *hb, const int item, const int nv) {
            const int boxoff = item & 0x1ff, x = (item >> 9) & 15, yy = (item >> 13) & 15, off = (item >> 17) & 255, zpl = (item >> 25) & 7;
            const bool valid = (item >> 28) & 1;
#pragma unroll
            for (int i = 0; i < EPT; ++i) {
                const int4 d = *reinterpret_cast<const int4 *>(hb + i * HSZ + boxoff);
                put_if(i < nv && valid, y + (d.x + d.y * x + d.z * yy + d.w * zpl), fo[i * FOS + off], RedT{});
            }
        };

        int ldI[EPT], ldS[EPT], ldG[EPT];  // planes being gathered
        int scI[EPT], scS[EPT], scG[EPT];  // planes being scattered

        // gather plane zp into ring slot `rs` (free). The lanes' copies arrive on full[rs] by themselves; the 33rd arrival is
        // the caller's.
        auto issue_plane = [&](const int rs, const unsigned char *hb, const int zp) {
            uint64_t *bar = full + rs;
            T *dst = ring + (size_t)rs * PLANE;
            if (zp == 0) {
                cp_async_wait<N - 2>();  // headers of this group were issued N commit groups ago
                __syncwarp();
                plane_set(hb, ldI, ldS, ldG);
            }
            if (dbg & 8) {  // experiment: no loads (the barrier protocol stays)
            } else if (zp == 0 || zp == L) {
#pragma unroll
                for (int i = 0; i < EPT; ++i) {
                    const int k = ldG[i] & (AL - 1);
                    cp_async_value(dst + (sk_off0 + ws_slot_shift<T>(i) + Lay::DK(i) * k), u + boundary_id(hb, i, sk_x, sk_y, sk_cxy, zp));
                    cp_async_value(dst + (i * PS + k + tl), u + boundary_id(hb, i, x0, y0, 4, zp));
                    copy_if(in_second, dst + (i * PS + k + q1), u + boundary_id(hb, i, x1, y1, 4, zp));
                    asm volatile("" ::: "memory");  // one slot at a time: small register budget
                }
            } else {
                if (tl == 0) {
                    mbar_expect_tx(bar, TX_BYTES);
#pragma unroll
                    for (int i = 0; i < EPT; ++i) tma_load_1d(dst + i * PS, &tmap_u, ldG[i] & ~(AL - 1), bar);
                }
#pragma unroll
                for (int i = 0; i < EPT; ++i) {
                    const int k = ldG[i] & (AL - 1);
                    cp_async_value(dst + (sk_off0 + ws_slot_shift<T>(i) + Lay::DK(i) * k), u + ldI[i]);
                }
            }
            cp_async_mbar_arrive_noinc(bar);
#pragma unroll
            for (int i = 0; i < EPT; ++i) {
                ldG[i] += NI;
                ldI[i] += ldS[i];
            }
        };

        // output plane z of the group with header buffer hb, staged in slot ss
        auto scatter_plane = [&](const int ss, const unsigned char *hb, const int z, const int nv) {
            const T *src = stage + (size_t)ss * PLANE;
            if (z == 0 || z == L) {
#pragma unroll
                for (int i = 0; i < EPT; ++i) {
                    const bool ok = i < nv;
                    const int k = scG[i] & (AL - 1);
                    put_if(ok, y + boundary_id(hb, i, sk_x, sk_y, sk_cxy, z), src[sk_off0 + ws_slot_shift<T>(i) + Lay::DK(i) * k], RedT{});
                    put_if(ok, y + boundary_id(hb, i, x0, y0, 4, z), src[i * PS + k + tl], RedT{});
                    put_if(ok && in_second, y + boundary_id(hb, i, x1, y1, 4, z), src[i * PS + k + q1], RedT{});
                    asm volatile("" ::: "memory");
                }
            } else {
                if constexpr (FACEOUT != 1) {
                    // FACEOUT 2: the lanes of the rows y = 0 / L with 0 < x < L only (the x = 0 / x = L positions are swept)
                    const bool mine = FACEOUT == 0 || (sk_x > 0 && sk_x < L);
#pragma unroll
                    for (int i = 0; i < EPT; ++i) {
                        const int k = scG[i] & (AL - 1);
                        put_if(i < nv && mine, y + scI[i], src[sk_off0 + ws_slot_shift<T>(i) + Lay::DK(i) * k], RedT{});
                    }
                }
#pragma unroll
                for (int i = 0; i < EPT; ++i) {
                    const int k = scG[i] & (AL - 1);
                    put_if(i < nv, y + (scG[i] + tl), src[i * PS + k + tl], PlainT{});
                    put_if(i < nv && in_second, y + (scG[i] + q1), src[i * PS + k + q1], PlainT{});
                }
            }
#pragma unroll
            for (int i = 0; i < EPT; ++i) {
                scG[i] += NI;
                scI[i] += scS[i];
            }
        };

        // distributed apply: the groups from dep.first_group on read ghost values; before the first plane of such a group is
        // gathered, wait until the receiver half of the halo gather (another stream, another SM) has announced the ghost block
        // of x. Bounded (2 s): a flag that never comes must not hang the GPU -- the apply then runs on stale ghosts and the
        // communicator's own time-out reports the failed exchange.
        auto wait_ghosts = [&](const int g_local) {
            if (dep.flag == nullptr || group_of(g_local) < dep.first_group) return;
            unsigned long long v;
            asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(dep.flag) : "memory");
            if (v < dep.want) {
                unsigned long long t0, t1;
                asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
                do {
                    __nanosleep(200);
                    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(dep.flag) : "memory");
                    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t1));
                } while (v < dep.want && t1 - t0 < 2000000000ull);
            }
            asm volatile("fence.proxy.async;" ::: "memory");  // the TMA reads of the ghost block come after the acquire
            __syncwarp();
        };
        // ---- prologue: headers of groups 0 and 1, planes 0 .. D-1
        issue_headers(0);
        cp_async_commit();
        cp_async_wait<0>();
        __syncwarp();
        issue_headers(1);
        wait_ghosts(0);
#pragma unroll 1
        for (int s = 0; s < D; ++s) {
            issue_plane(s, hdr_bytes, s);  // D < N <= T_steps: all in group 0
            cp_async_commit();
            if (tl == 0) mbar_arrive(full + s);
        }
        // ---- stream: iteration t follows the compute warp's step t (t = T_steps: the final emit).
        // (gn, zn, hn) = group, plane, header buffer of plane t + D; rs = t % D = its ring slot = the staging slot of output
        // plane t - 1; ph = parity of sfull[rs]; (zs, hs, gs) = plane, header buffer, group of output plane t - 1.
        int gn = 0, zn = D, hn = 0, rs = 0;
        uint32_t ph = 0;
        int zs = 0, hs = 0, gs = 0, nv = EPT;
#pragma unroll 1
        for (int t = 0; t <= T_steps; ++t) {
            mbar_wait_hint<HINT>(sfull + rs, ph);
            const bool more = t + D < T_steps;
            const bool closing = more || (t + D == T_steps);  // plane T_steps is a phantom: it only carries "staging slot free"
            if (more) {
                if (zn == 0) {
                    issue_headers(gn + 1);  // headers of group g+2 travel with the first plane of g+1
                    wait_ghosts(gn);
                }
                issue_plane(rs, hdr_bytes + hn * HBUF, zn);
            } else if (closing) {
                cp_async_mbar_arrive_noinc(full + rs);
            }
            cp_async_commit();
            if (++zn == N) { zn = 0; ++gn; hn = (hn == 2) ? 0 : hn + 1; }
            if (t > 0) {
                const unsigned char *hb = hdr_bytes + hs * HBUF;
                if (zs == 0) {
                    plane_set(hb, scI, scS, scG);
                    const int left = (int)ne - group_of(gs) * EPT;
                    nv = left < EPT ? left : EPT;
                }
                if (!(dbg & 4)) scatter_plane(rs, hb, zs, nv);
            }
            if (tl == 0) {
                if (closing) mbar_arrive(full + rs);
            }
            if (t > 0) {
                if constexpr (FACEOUT) {
                    // after the closing arrival (the sweeps read the lateral buffer, not the staging slot; the half they read is
                    // rewritten 6 steps later at the earliest, behind the closing arrivals of the iterations in between)
                    if (!(dbg & 4)) {
                        const unsigned char *hb = hdr_bytes + hs * HBUF;
                        if constexpr (FACEOUT == 2) {
                            // one sweep instruction set per iteration (planes 1 .. 4 are complete from zs = 4 on, planes 5 .. L-1
                            // from zs = L-1 on): an iteration that did a whole half at once fell behind its compute warp
                            if (zs >= 4) sweep(hb, zs == 4 ? sweepA[0] : (zs == 5 ? sweepA[1] : (zs == 6 ? sweepA[2] : (zs == 7 ? sweepB[0] : sweepB[1]))), nv);
                        } else if (zs == 4) {
#pragma unroll
                            for (int q = 0; q < NSA; ++q) sweep(hb, sweepA[q], nv);
                        } else if (zs == L - 1) {
#pragma unroll
                            for (int q = 0; q < NSB; ++q) sweep(hb, sweepB[q], nv);
                        }
                    }
                }
                if (++zs == N) { zs = 0; ++gs; hs = (hs == 2) ? 0 : hs + 1; }
            }
            if (++rs == D) { rs = 0; ph ^= 1u; }
        }
        cp_async_wait<0>();
    } else {
        // ========================================================================================= compute warp
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;\n" ::"n"(REG_COMPUTE));
        if (my_groups == 0) return;
        const int cl = tl < EPT * N ? tl : EPT * N - 1;
        const int slot = cl / N;
        const int yrow = cl - slot * N;
        int a0[3], am[3];
#pragma unroll
        for (int iy = 0; iy < 3; ++iy) {
            const int yy = yrow + iy - 1;  // unclamped, see ws_offset
            a0[iy] = ws_offset<T, L>(slot, 0, yy);
            am[iy] = ws_offset<T, L>(slot, 1, yy);
        }
        bool rint[3];
#pragma unroll
        for (int iy = 0; iy < 3; ++iy) {
            const int yy = yrow + iy - 1;
            rint[iy] = (yy >= 1 && yy <= L - 1);
        }
        const int dk = (slot == 0) ? Lay::DK(0) : (slot == 1 ? Lay::DK(1) : Lay::DK(2));
        int Bs = 0, Bs_prev = 0;  // id of lattice node (1, 1, 0) of this lane's element, and of the previous group's
        Sshex8LaplacianRow<T, L> row;
        row.clear();

        // FACEOUT: positions of this lane's x = 0 / x = L nodes among the 32 skeleton positions of a plane (the rows y = 0 / L
        // are positions 0 .. L and N .. N + L, the columns follow); x = 1 .. L-1 of a lateral row go to positions posA + x
        const int posA = FACEOUT == 2 ? (yrow == 0 ? 0 : (yrow == L ? 2 : 3 + yrow)) : (yrow == 0 ? 0 : (yrow == L ? N : 2 * N + yrow - 1));
        const int posB = FACEOUT == 2 ? (yrow == 0 ? 1 : (yrow == L ? 3 : 10 + yrow))
                                      : (yrow == 0 ? L : (yrow == L ? N + L : 2 * N + (L - 1) + yrow - 1));
        T *fo_lane = fo + slot * FOS;
        const int mid_stride = rint[1] ? 1 : (L - 1);
        // this lane's row of output plane z of the element whose interior starts at B, into staging slot ss
        auto stage_row = [&](const int ss, const int z, const int B, const T(&vals)[N]) {
            T *dst = stage + (size_t)ss * PLANE;
            const int k = (B + NI * z) & (AL - 1);
            if constexpr (FACEOUT == 1) {
                if (z != 0 && z != L) {  // warp-uniform
                    T *f = fo_lane + (z - 1);
                    f[posA * (L - 1)] = vals[0];
                    f[posB * (L - 1)] = vals[L];
                    T *dmid = rint[1] ? (dst + k + am[1]) : (f + (posA + 1) * (L - 1));
#pragma unroll
                    for (int x = 1; x < L; ++x) dmid[(x - 1) * mid_stride] = vals[x];
                    return;
                }
            }
            if constexpr (FACEOUT == 2) {
                if (z != 0 && z != L) {  // warp-uniform: x = 0 / x = L into the lateral buffer, the rest staged as always
                    T *f = fo_lane + (z - 1);
                    f[posA * (L - 1)] = vals[0];
                    f[posB * (L - 1)] = vals[L];
                    T *dm2 = (rint[1] ? dst + k : dst + dk * k) + am[1];
#pragma unroll
                    for (int x = 1; x < L; ++x) dm2[x - 1] = vals[x];
                    return;
                }
            }
            T *ds = dst + dk * k;
            T *dm = (rint[1] ? dst + k : ds) + am[1];
            ds[a0[1]] = vals[0];
#pragma unroll
            for (int x = 1; x < L; ++x) dm[x - 1] = vals[x];
            ds[a0[1] + (rint[1] ? L - 1 : L)] = vals[L];
        };

        if constexpr (PHASED) {
            // Static accumulator phases (plane z accumulates in acc[z % 3]: no register rotation), four step bodies: z = 0,
            // and the interior class in phases 0 / 1 / 2; z = L runs through the phase-2 body with its accumulator pre-doubled,
            // the edge terms applied in one piece afterwards, and the (meaningless) contribution to plane L + 1 zeroed.
            int rs = 0;
            uint32_t ph = 0;
            bool first = true;
            using I0 = std::integral_constant<int, 0>;
            using I1 = std::integral_constant<int, 1>;
            using I2 = std::integral_constant<int, 2>;
            auto z_step = [&](auto zc_tag, auto p_tag, const int g_local, const int zp) {
                constexpr int ZC = decltype(zc_tag)::value, P = decltype(p_tag)::value;
                const T *plane = ring + (size_t)rs * PLANE;
                mbar_wait_hint<HINT>(full + rs, ph);
                if constexpr (ZC == 0) {
                    const Sshex8Header *h = reinterpret_cast<const Sshex8Header *>(hdr_bytes + (g_local % 3) * HBUF) + slot;
                    T f[6];
#pragma unroll
                    for (int k = 0; k < 6; ++k) f[k] = (T)h->fff[k];
                    row.set(f, yrow);
                    Bs_prev = Bs;
                    Bs = h->box[13][0] + h->box[13][1] + h->box[13][2];
                }
                const int kin = (Bs + NI * zp) & (AL - 1);
                const T *pk = plane + kin, *ps = plane + dk * kin;
                auto loader = [&](const int iy, T(&r)[N]) {
                    const T *pm = (rint[iy] ? pk : ps) + am[iy];
                    r[0] = ps[a0[iy]];
#pragma unroll
                    for (int x = 1; x < L; ++x) r[x] = pm[x - 1];
                    r[L] = ps[a0[iy] + (rint[iy] ? L - 1 : L)];
                };
                const bool zhi = (P == 2) && zp == L;  // warp-uniform
                if (P == 2 && zhi) {
#pragma unroll
                    for (int x = 0; x < N; ++x) row.acc[2][x] *= T(2);
                }
                if (!(dbg & 2)) row.template step_class_with<ZC, P>(loader);
                if (P == 2 && zhi) {
                    T r0[3], rL[3], rmid[N];
#pragma unroll
                    for (int iy = 0; iy < 3; ++iy) {
                        r0[iy] = ps[a0[iy]];
                        rL[iy] = ps[a0[iy] + (rint[iy] ? L - 1 : L)];
                    }
                    loader(1, rmid);
                    row.z_boundary_terms(row.acc[2], T(2), r0, rL, rmid);
#pragma unroll
                    for (int x = 0; x < N; ++x) {
                        row.acc[2][x] *= T(0.5);
                        row.acc[0][x] = T(0);  // what the interior-class code sent towards plane L + 1
                    }
                }
                // output plane zp - 1 (z = 0: the previous group's plane L) is final: stage it, hand the slot over, recycle
                if (!first) stage_row(rs, ZC == 0 ? L : zp - 1, ZC == 0 ? Bs_prev : Bs, row.acc[(P + 2) % 3]);
                first = false;
                mbar_arrive(sfull + rs);
#pragma unroll
                for (int x = 0; x < N; ++x) row.acc[(P + 2) % 3][x] = T(0);
                if (++rs == D) { rs = 0; ph ^= 1u; }
            };
            static_assert(N % 3 == 0 && L % 3 == 2, "phase pattern: N a multiple of 3, plane L in phase 2");
#pragma unroll 1
            for (int g_local = 0; g_local < my_groups; ++g_local) {
#pragma unroll 1
                for (int zz = 0; zz < N; zz += 3) {
                    if (zz == 0)
                        z_step(I0{}, I0{}, g_local, 0);
                    else
                        z_step(I1{}, I0{}, g_local, zz);
                    z_step(I1{}, I1{}, g_local, zz + 1);
                    z_step(I1{}, I2{}, g_local, zz + 2);
                }
            }
            mbar_wait_hint<HINT>(full + rs, ph);  // the phantom plane: staging slot rs has been scattered
            stage_row(rs, L, Bs, row.acc[2]);
            mbar_arrive(sfull + rs);
            return;
        }
        int g_local = 0, zp = 0, rs = 0;
        uint32_t ph = 0;
#pragma unroll 1
        for (int t = 0; t < T_steps; ++t) {
            const T *plane = ring + (size_t)rs * PLANE;
            mbar_wait_hint<HINT>(full + rs, ph);
            if (zp == 0) {
                const Sshex8Header *h = reinterpret_cast<const Sshex8Header *>(hdr_bytes + (g_local % 3) * HBUF) + slot;
                T f[6];
#pragma unroll
                for (int k = 0; k < 6; ++k) f[k] = (T)h->fff[k];
                row.set(f, yrow);  // acc[0] still holds the last output plane of the previous group (ZC = 0 leaves it alone)
                Bs_prev = Bs;
                Bs = h->box[13][0] + h->box[13][1] + h->box[13][2];
            }
            const int kin = (Bs + NI * zp) & (AL - 1);
            const T *pk = plane + kin, *ps = plane + dk * kin;
            if (!(dbg & 2)) row.step_with(zp, [&](const int iy, T(&r)[N]) {
                const T *pm = (rint[iy] ? pk : ps) + am[iy];
                r[0] = ps[a0[iy]];
#pragma unroll
                for (int x = 1; x < L; ++x) r[x] = pm[x - 1];
                r[L] = ps[a0[iy] + (rint[iy] ? L - 1 : L)];
            });
            if (t > 0) stage_row(rs, zp == 0 ? L : zp - 1, zp == 0 ? Bs_prev : Bs, row.acc[0]);
            mbar_arrive(sfull + rs);
            row.rotate();
            if (++zp == N) { zp = 0; ++g_local; }
            if (++rs == D) { rs = 0; ph ^= 1u; }
        }
        mbar_wait_hint<HINT>(full + rs, ph);  // the phantom plane: staging slot rs has been scattered
        stage_row(rs, L, Bs, row.acc[0]);
        mbar_arrive(sfull + rs);
    }
}


}  // namespace b200fem
