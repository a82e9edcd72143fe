// sshex8 Laplacian, "plane-streaming" kernel (boxes connectivity): the production path of the north-star
// operator. See sshex8_laplacian.cu for what is computed and sshex8_laplacian_core.cuh for the per-thread
// arithmetic; this file is only about how the data moves.
//
// Shape. A persistent grid of one CTA per SM. A CTA is NTEAM independent *teams* of TW warps; a team owns
// EPT = floor(32 TW / N) macro elements at a time (N = L+1 lattice rows per element, one lane per row) and
// walks an endless stream of z-planes:  step t  ->  group t / N (EPT consecutive elements), plane t % N.
// Per step a team
//   1. computes plane t out of its shared-memory ring slot t % D (Sshex8LaplacianRow::step: 27 LDS.64 and
//      ~250 DFMA per lane), and writes the output plane that became final to a staging buffer;
//   2. waits for its own cp.async copies of plane t+1, then passes the ONE team barrier of the step
//      (__syncwarp for TW = 1, a named barrier otherwise);
//   3. issues the gather of plane t+D straight into the ring slot it just finished with
//      (cp.async 8/4 B per node, addresses from the element's 27 box descriptors: no registers held,
//      (D-1) planes = (D-1) EPT N^2 sizeof(T) bytes in flight per team);
//   4. scatters the staged output plane: interior nodes (owned by exactly one element) with plain
//      coalesced stores when overwriting, skeleton nodes with RED.ADD.
// There is no CTA-wide barrier and no phase in which the memory system or the FP64 pipe waits for the
// other: loads, DFMAs and stores of different planes overlap inside every team, and the 8 warps of an SM
// drift apart freely.
//
// Element headers. The plan stores, per element, 27 box descriptors (16 B each) + the 6 fp32 FFF + pad =
// 29 x 16 B, fetched with cp.async one whole group ahead of their first use (three header buffers per team).
#pragma once
#include <type_traits>

#include "common.cuh"
#include "sshex8_laplacian_core.cuh"

namespace b200fem {

struct Sshex8Header {   // 464 B, 16-byte aligned
    int32_t box[27][4];  // base, sx, sy, sz  (id = base + sx x + sy y + sz z)
    float fff[6];
    float pad[2];
};
static_assert(sizeof(Sshex8Header) == 29 * 16, "header must be 29 x 16 B");

// Warp-specialised level-8 kernel (sshex8_laplacian_ws.cu / .cuh): `headers` must be readable up to the next multiple
// of 3 elements (padding = copies of the last element; only valid slots are scattered); compact numbering
// (interior block of an element = one contiguous id run, x fastest) and 16-byte aligned x required. real_type is
// B200FEM_F64 / B200FEM_F32; `plain` = interior nodes are stored, not added (overwriting apply).
// elements of groups >= first_group wait in the kernel until *flag >= want (acquire, GPU scope): see b200fem_op_apply_overwrite_after
struct WsDependency {
    const unsigned long long *flag = nullptr;
    unsigned long long want = 0;
    int first_group = 0;
};
int sshex8_laplacian_ws_launch(int real_type, bool plain, int cfg, ptrdiff_t ne, ptrdiff_t nn, const Sshex8Header *headers, const void *x,
                               void *y, int max_ctas, cudaStream_t s, const WsDependency *dep = nullptr);

// The 4/8-byte copies and the scatter stores below carry NO "memory" clobber on purpose: their ordering
// against ordinary shared-memory traffic is established by the team barrier and cp.async.wait_group (which
// do clobber), and without it the compiler is free to hoist the descriptor loads of all items of a plane
// above the first copy instead of serialising LDS -> IMAD chain -> LDGSTS item by item.
__device__ __forceinline__ void cp_async_4(void *smem, const void *gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem));
}
__device__ __forceinline__ void cp_async_8(void *smem, const void *gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem));
}
__device__ __forceinline__ void cp_async_16(void *smem, const void *gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int K>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(K) : "memory");
}
template <typename T>
__device__ __forceinline__ void cp_async_value(T *smem, const T *gmem) {
    if (sizeof(T) == 8)
        cp_async_8(smem, gmem);
    else
        cp_async_4(smem, gmem);
}

template <int Z0, int Z1, class F>
__device__ __forceinline__ void static_for_z(F &&f) {
    if constexpr (Z0 < Z1) {
        f(std::integral_constant<int, Z0>{});
        static_for_z<Z0 + 1, Z1>(f);
    }
}

template <int TW>
__device__ __forceinline__ void team_barrier(int team) {
    if (TW == 1) {
        __syncwarp();
    } else {
        asm volatile("bar.sync %0, %1;\n" ::"r"(team + 1), "n"(TW * 32) : "memory");
    }
}

template <typename T, int L, int TW, int NTEAM, int D>
struct Sshex8StreamCfg {
    static constexpr int N = L + 1;
    static constexpr int NN = N * N;
    static constexpr int TL = TW * 32;               // lanes per team
    static constexpr int EPT = TL / N;               // elements per team-group
    static constexpr int PLANE = EPT * NN;           // values per plane of a group
    static constexpr int KMAX = (PLANE + TL - 1) / TL;  // copy/scatter items per lane per plane
    static constexpr int HDR16 = EPT * 29;           // 16-byte chunks of a group's headers
    static constexpr size_t ring_bytes = (size_t)D * PLANE * sizeof(T);
    static constexpr size_t stage_bytes = (size_t)3 * PLANE * sizeof(T);
    static constexpr size_t hdr_bytes = (size_t)3 * EPT * sizeof(Sshex8Header);
    static constexpr size_t team_bytes = ((ring_bytes + stage_bytes + 15) / 16) * 16 + hdr_bytes;
    static constexpr size_t smem_bytes = team_bytes * NTEAM;
    static constexpr int threads = NTEAM * TL;
    static_assert(D >= 2 && D <= N, "ring depth must be in [2, N]");
    static_assert(EPT >= 1, "team too small for this level");
};

// interior-first order of the N x N positions of a lattice plane: the (L-1)^2 interior positions
// (contiguous in global memory for the compact numbering), then the y = 0 row, the y = L row, the x = 0
// column and the x = L column
template <int L>
__device__ __forceinline__ void plane_pos_to_xy(const int q, int &x, int &yy) {
    constexpr int N = L + 1, NI = (L - 1) * (L - 1);
    if (q < NI) {
        yy = 1 + q / (L - 1);
        x = 1 + q - (yy - 1) * (L - 1);
    } else if (q < NI + N) {
        yy = 0;
        x = q - NI;
    } else if (q < NI + 2 * N) {
        yy = L;
        x = q - NI - N;
    } else if (q < NI + 2 * N + (L - 1)) {
        x = 0;
        yy = 1 + q - (NI + 2 * N);
    } else {
        x = L;
        yy = 1 + q - (NI + 2 * N + (L - 1));
    }
}

// one value: plain streaming store (exclusive interior node, y overwritten) or RED.ADD, branch-free
__device__ __forceinline__ void store_or_red(double *p, const double v, const int plain) {
    asm volatile(
        "{\n .reg .pred q;\n setp.ne.s32 q, %2, 0;\n @q st.global.cs.f64 [%0], %1;\n @!q red.global.add.f64 [%0], %1;\n}\n" ::"l"(p),
        "d"(v), "r"(plain));
}
__device__ __forceinline__ void store_or_red(float *p, const float v, const int plain) {
    asm volatile(
        "{\n .reg .pred q;\n setp.ne.s32 q, %2, 0;\n @q st.global.cs.f32 [%0], %1;\n @!q red.global.add.f32 [%0], %1;\n}\n" ::"l"(p),
        "f"(v), "r"(plain));
}

// Per-lane description of one copy/scatter item whose (slot, position) does not change over time.
struct PlaneItem {
    int x, yy;   // lattice coordinates in the plane
    int hoff;    // byte offset of its box descriptor inside a header buffer, cz = 0 (slot*464 + cxy*16)
    int soff;    // element offset inside a plane buffer (slot*NN + yy*N + x)
    int center;  // 1 if the position is interior in x and y
    int slot;    // element slot of the group, -1 = no item
};

// PLAIN_INTERIOR: interior nodes are exclusive to their element and y is being overwritten -> plain stores.
// CHECK_TAIL: the last group may hold fewer than EPT elements (only the small clean-up launch pays for the
// checks; the main launch covers floor(ne / EPT) full groups).
template <typename T, int L, int TW, int NTEAM, int D, bool PLAIN_INTERIOR, bool CHECK_TAIL>
__global__ void __launch_bounds__(NTEAM *TW * 32, 1)
    sshex8_laplacian_stream_kernel(const ptrdiff_t ne, const Sshex8Header *__restrict__ headers,
                                   const T *__restrict__ u, T *__restrict__ y) {
    using Cfg = Sshex8StreamCfg<T, L, TW, NTEAM, D>;
    constexpr int N = Cfg::N, NN = Cfg::NN, TL = Cfg::TL, EPT = Cfg::EPT, PLANE = Cfg::PLANE;
    // Item mapping. If a team has no more lanes than a plane has positions (TL <= NN), lane `tl` serves
    // position r*TL + tl of EVERY slot for the R full rounds r (so x, y, box class and smem offset are
    // per-lane constants and the slot enters as an immediate), and the LO = NN - R*TL left-over
    // positions of all slots are dealt round-robin over XR extra rounds. Otherwise (small levels) all
    // PLANE positions are dealt round-robin.
    constexpr bool STRUCT = (TL <= NN);
    constexpr int R = STRUCT ? NN / TL : 0;
    constexpr int LO = STRUCT ? NN - R * TL : NN;
    constexpr int XR = (LO * EPT + TL - 1) / TL;
    constexpr int NIT = EPT * R + XR;  // items per lane per plane
    constexpr int HSZ = (int)sizeof(Sshex8Header);

    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int team = threadIdx.x / TL;
    const int tl = threadIdx.x - team * TL;  // lane within the team
    unsigned char *tbase = smem_raw + (size_t)team * Cfg::team_bytes;
    T *ring = reinterpret_cast<T *>(tbase);                       // [D][PLANE]
    T *stage = ring + (size_t)D * PLANE;                          // [3][PLANE]
    unsigned char *hdr_bytes = tbase + ((Cfg::ring_bytes + Cfg::stage_bytes + 15) / 16) * 16;  // [3][EPT] headers
    constexpr int HBUF = EPT * HSZ;

    // ---- work of this team: groups gteam, gteam + nteams, ...
    const ptrdiff_t ngroups = (ne + EPT - 1) / EPT;
    const ptrdiff_t nteams = (ptrdiff_t)gridDim.x * NTEAM;
    const ptrdiff_t gteam = (ptrdiff_t)blockIdx.x * NTEAM + team;
    if (gteam >= ngroups) return;
    const int my_groups = (int)((ngroups - gteam + nteams - 1) / nteams);
    const int T_steps = my_groups * N;

    // ---- per-lane items: (x, y) in the plane, header byte offset of the box class at cz = 0, plane offset
    int it_x[NIT], it_y[NIT], it_h[NIT], it_s[NIT];
    unsigned center_mask = 0, valid_mask = 0;
    int it_slot[XR > 0 ? XR : 1];
#pragma unroll
    for (int i = 0; i < NIT; ++i) {
        int slot, pos;
        bool valid = true;
        if (i < EPT * R) {
            slot = i / (R > 0 ? R : 1);
            pos = (i - slot * R) * TL + tl;
        } else {
            const int j = (i - EPT * R) * TL + tl;
            valid = j < LO * EPT;
            slot = valid ? j / LO : 0;
            pos = valid ? R * TL + (j - slot * LO) : 0;
            it_slot[i - EPT * R] = valid ? slot : EPT;  // EPT = never valid
        }
        int x, yy;
        plane_pos_to_xy<L>(pos, x, yy);
        const int cx = (x == 0) ? 0 : (x == L ? 2 : 1), cy = (yy == 0) ? 0 : (yy == L ? 2 : 1);
        it_x[i] = x;
        it_y[i] = yy;
        it_h[i] = slot * HSZ + (cx + 3 * cy) * 16;
        it_s[i] = slot * NN + yy * N + x;
        if (cx == 1 && cy == 1) center_mask |= 1u << i;
        if (valid) valid_mask |= 1u << i;
    }

    auto group_of = [&](int g_local) -> ptrdiff_t { return gteam + (ptrdiff_t)g_local * nteams; };
    auto valid_slots = [&](int g_local) -> int {
        if (!CHECK_TAIL) return EPT;
        const ptrdiff_t left = ne - group_of(g_local) * EPT;
        return (int)(left < (ptrdiff_t)EPT ? left : (ptrdiff_t)EPT);
    };
    // which items exist for this group (bit i)
    auto item_mask = [&](int g_local) -> unsigned {
        if (!CHECK_TAIL) return valid_mask;
        const int nvalid = valid_slots(g_local);
        unsigned m = 0;
#pragma unroll
        for (int i = 0; i < NIT; ++i) {
            const int slot = (i < EPT * R) ? i / (R > 0 ? R : 1) : it_slot[i - EPT * R];
            if (slot < nvalid) m |= 1u << i;
        }
        return m & valid_mask;
    };

    auto issue_headers = [&](int g_local) {
        if (g_local >= my_groups) return;
        const int nvalid = valid_slots(g_local);
        const int4 *src = reinterpret_cast<const int4 *>(headers + group_of(g_local) * EPT);
        int4 *dst = reinterpret_cast<int4 *>(hdr_bytes + (g_local % 3) * HBUF);
        for (int i = tl; i < nvalid * 29; i += TL) cp_async_16(dst + i, src + i);
    };

    // Node ids of a plane: id(z) = I + S z with (I, S) fixed over the z-interior planes of a group, so the
    // 7 interior planes of a level-8 element cost one IMAD per node; planes 0 and L read their own boxes.
    auto plane_set = [&](int g_local, int zc, int (&I)[NIT], int (&S)[NIT]) {
        const unsigned char *hb = hdr_bytes + (g_local % 3) * HBUF + zc * 9 * 16;
#pragma unroll
        for (int i = 0; i < NIT; ++i) {
            const int4 d = *reinterpret_cast<const int4 *>(hb + it_h[i]);
            I[i] = d.x + d.y * it_x[i] + d.z * it_y[i];
            S[i] = d.w;
        }
    };

    int ldI[NIT], ldS[NIT];  // z-interior set of the group whose planes are being gathered
    int scI[NIT], scS[NIT];  // z-interior set of the group whose planes are being scattered

    auto issue_plane = [&](int s) {  // gather plane of step s into ring slot s % D
        if (s >= T_steps) return;
        const int g_local = s / N, zp = s - g_local * N;
        const unsigned m = item_mask(g_local);
        T *dst = ring + (size_t)(s % D) * PLANE;
        if (zp == 0) plane_set(g_local, 1, ldI, ldS);  // the set of the group the gather is entering
        int id[NIT];
        if (zp == 0 || zp == L) {
            int I[NIT], S[NIT];
            plane_set(g_local, zp == 0 ? 0 : 2, I, S);
#pragma unroll
            for (int i = 0; i < NIT; ++i) id[i] = I[i] + S[i] * zp;
        } else {
#pragma unroll
            for (int i = 0; i < NIT; ++i) id[i] = ldI[i] + ldS[i] * zp;
        }
#pragma unroll
        for (int i = 0; i < NIT; ++i) {
            if ((m >> i) & 1u) cp_async_value(dst + it_s[i], u + id[i]);
        }
    };

    auto scatter_plane = [&](const T *src, int g_local, int z) {
        const unsigned m = item_mask(g_local);
        int id[NIT];
        unsigned plain = 0;
        if (z == 0 || z == L) {
            int I[NIT], S[NIT];
            plane_set(g_local, z == 0 ? 0 : 2, I, S);
#pragma unroll
            for (int i = 0; i < NIT; ++i) id[i] = I[i] + S[i] * z;
        } else {
#pragma unroll
            for (int i = 0; i < NIT; ++i) id[i] = scI[i] + scS[i] * z;
            if (PLAIN_INTERIOR) plain = center_mask;
        }
        T val[NIT];
#pragma unroll
        for (int i = 0; i < NIT; ++i) val[i] = src[it_s[i]];
#pragma unroll
        for (int i = 0; i < NIT; ++i) {
            if ((m >> i) & 1u) store_or_red(y + id[i], val[i], (int)((plain >> i) & 1u));
        }
    };

    // ---- prologue: headers of groups 0 and 1, then planes 0 .. D-1
    issue_headers(0);
    cp_async_commit();
    cp_async_wait<0>();
    team_barrier<TW>(team);
    issue_headers(1);
#pragma unroll 1
    for (int s = 0; s < D; ++s) {
        issue_plane(s);
        cp_async_commit();
    }
    cp_async_wait<D - 1>();
    team_barrier<TW>(team);

    // ---- stream
    // Lanes beyond EPT * N (5 of 32 at L = 8) shadow the last row: they run the same arithmetic on the same shared-memory
    // addresses and store the same values, so the compute section needs no divergent region (the FP64 pipe charges a
    // warp instruction the same for 27 or 32 active lanes; reads of one address broadcast).
    const int cl = tl < EPT * N ? tl : EPT * N - 1;
    const bool compute_lane = true;
    const int slot = cl / N;
    const int yrow = cl - slot * N;
    Sshex8LaplacianRow<T, L> row;

    // Two forms of the same loop. fp32: z unrolled (static accumulator phases of the core: no register rotation, no
    // per-step branching; 0.82 -> 0.71 ms at C3). fp64: the rolled loop -- unrolled, the 27 fp64 accumulators +
    // 27 coefficients no longer fit 255 registers (spills, 0.98 -> 1.14 ms).
    if constexpr (sizeof(T) == 4) {
    // z is unrolled (static accumulator phases of the core: no register rotation, no per-step branching); ring and
    // staging slots are compile-time whenever the ring depth divides N
    constexpr bool RING_STATIC = (N % D == 0);
#pragma unroll 1
    for (int g_local = 0; g_local < my_groups; ++g_local) {
        static_for_z<0, N>([&](auto zc) {
            constexpr int zp = decltype(zc)::value;
            const int t = g_local * N + zp;
            const int rslot = RING_STATIC ? (zp % D) : (t % D);
            const int s0 = (N % 3 == 0) ? (zp % 3) : (t % 3), s1 = (N % 3 == 0) ? ((zp + 1) % 3) : ((t + 1) % 3);
            const T *plane = ring + (size_t)rslot * PLANE + slot * NN;
            T *st0 = stage + (size_t)s0 * PLANE + slot * NN;
            T *st1 = stage + (size_t)s1 * PLANE + slot * NN;
            if constexpr (zp == 0) plane_set(g_local, 1, scI, scS);
            if (compute_lane) {
                if constexpr (zp == 0) {
                    const Sshex8Header *h = reinterpret_cast<const Sshex8Header *>(hdr_bytes + (g_local % 3) * HBUF) + slot;
                    T f[6];
#pragma unroll
                    for (int k = 0; k < 6; ++k) f[k] = (T)h->fff[k];
                    row.init(f, yrow);
                }
                row.template step_static<zp>([plane, yrow](const int iy, T(&r)[N]) {
                    int yy = yrow + iy - 1;
                    yy = yy < 0 ? 0 : (yy > L ? L : yy);
                    const T *rp = plane + yy * N;
#pragma unroll
                    for (int x = 0; x < N; ++x) r[x] = rp[x];
                });
                if constexpr (zp > 0) {  // output plane zp-1 is final
                    T *orow = st0 + yrow * N;
                    const T(&fin)[N] = row.template finished_plane<zp>();
#pragma unroll
                    for (int x = 0; x < N; ++x) orow[x] = fin[x];
                    row.template recycle<zp>();
                }
                if constexpr (zp == L) {
                    T *orow = st1 + yrow * N;
                    const T(&fin)[N] = row.template finished_plane<L + 1>();
#pragma unroll
                    for (int x = 0; x < N; ++x) orow[x] = fin[x];
                }
            }
            cp_async_wait<D - 2>();
            team_barrier<TW>(team);

            // headers of group g+2 travel with the first plane of group g+1
            const int s_next = t + D;
            if (s_next < T_steps && (s_next % N) == 0) issue_headers(s_next / N + 1);
            issue_plane(s_next);
            cp_async_commit();

            if constexpr (zp > 0) scatter_plane(stage + (size_t)s0 * PLANE, g_local, zp - 1);
            if constexpr (zp == L) scatter_plane(stage + (size_t)s1 * PLANE, g_local, L);
        });
    }
    } else {
    int g_local = 0, zp = 0;
#pragma unroll 1
    for (int t = 0; t < T_steps; ++t) {
        const T *plane = ring + (size_t)(t % D) * PLANE + slot * NN;
        T *st0 = stage + (size_t)(t % 3) * PLANE + slot * NN;
        T *st1 = stage + (size_t)((t + 1) % 3) * PLANE + slot * NN;
        if (zp == 0) plane_set(g_local, 1, scI, scS);
        if (compute_lane) {
            if (zp == 0) {
                const Sshex8Header *h = reinterpret_cast<const Sshex8Header *>(hdr_bytes + (g_local % 3) * HBUF) + slot;
                T f[6];
#pragma unroll
                for (int k = 0; k < 6; ++k) f[k] = (T)h->fff[k];
                row.init(f, yrow);
            }
            row.step(zp, plane, N);
            row.write_and_rotate(zp, st0, N);  // output plane zp-1 (nothing at zp == 0)
            if (zp == L) row.write_last(st1, N);
        }
        cp_async_wait<D - 2>();
        team_barrier<TW>(team);

        // headers of group g+2 travel with the first plane of group g+1
        const int s_next = t + D;
        if (s_next < T_steps && (s_next % N) == 0) issue_headers(s_next / N + 1);
        issue_plane(s_next);
        cp_async_commit();

        if (zp > 0) scatter_plane(stage + (size_t)(t % 3) * PLANE, g_local, zp - 1);
        if (zp == L) scatter_plane(stage + (size_t)((t + 1) % 3) * PLANE, g_local, L);

        if (++zp == N) {
            zp = 0;
            ++g_local;
        }
    }
    }
    cp_async_wait<0>();
}

}  // namespace b200fem
