// sshex8 Laplacian, "entity-box TMA" kernel: the production path for compact numberings at L >= 4.
//
// What is computed: sshex8_laplacian_core.cuh (same per-lane arithmetic as the plane-streaming kernel).
// How the data moves. Every one of the 27 entity boxes of a macro element (1 interior block, 6 faces, 12 edges,
// 8 corners) is a CONTIGUOUS run of node ids for the reference's numbering, so an element's (L+1)^3 input values
// are 27 one-dimensional bulk copies (cp.async.bulk, the TMA engine) instead of (L+1)^3 per-node 8-byte copies,
// and its results leave the same way: cp.async.bulk shared->global for the exclusive interior block when y is being
// overwritten, cp.reduce.async.bulk .add for everything that other elements contribute to. The LSU / L1 pipe,
// which limited the streaming kernel (ncu: L1/TEX throughput 70 %, 8 cycles per LDGSTS), only serves the
// shared-memory reads and writes of the arithmetic now.
//
// Layout. A warp owns EPW = floor(32 / (L+1)) elements at a time (lane = (slot, y row), 27 of 32 lanes at L = 8)
// and one shared-memory buffer per slot in BOX layout: box b lives in a fixed region R_b of the buffer, shifted by
// `shift = gmin & (AL-1)` elements so that global and shared addresses agree modulo 16 bytes (AL = 16 / sizeof(T);
// bulk copies need 16-byte alignment on both sides and a multiple of 16 bytes). The copy of a box therefore covers
// [gmin - shift, gmin - shift + len), len = round_up(shift + count, AL): up to AL-1 neighbouring values ride along on
// either side; they are zeroed before the reduce so that they add 0.0 to their owners.
// Lattice node (x, y, z) of class c = (cx, cy, cz) sits at element offset  A_c + sx_c x + sy_c y + sz_c z  of the
// buffer (A_c = R_c + shift_c + base_c - gmin_c, precomputed per element at plan creation), i.e. the arithmetic
// reads its rows through three (address, stride) pairs per row, advanced by one add per z-step inside the element.
// Results are written IN PLACE (output plane z-1 after input plane z has been consumed; one __syncwarp per step).
//
// Pipeline. Warps are independent: each runs  [wait buffers free] -> issue 27 x EPW bulk loads on its own mbarrier
// -> wait -> 9 z-steps -> zero pads, fence.proxy.async -> issue 27 x EPW bulk stores / reduces -> next group.
// While a warp waits for its loads (~2 us) the other warp of its scheduler has the FP64 pipe to itself; element
// headers (352 B each) are prefetched one group ahead with one bulk copy per group.
#pragma once
#include <type_traits>

#include "common.cuh"
#include "sshex8_boxes.cuh"
#include "sshex8_laplacian_core.cuh"

namespace b200fem {

struct Sshex8TmaHeader {  // 352 B = 22 x 16
    int32_t gmin[27];     // smallest node id of box b
    int32_t pad0;
    int16_t desc[27][4];  // A, sx, sy, sz (element units, see above)
    float fff[6];
};
static_assert(sizeof(Sshex8TmaHeader) == 352, "header must be 22 x 16 B");

template <int L>
__host__ __device__ constexpr int tma_box_count(int b) {
    return ((b % 3) == 1 ? L - 1 : 1) * (((b / 3) % 3) == 1 ? L - 1 : 1) * ((b / 9) == 1 ? L - 1 : 1);
}
// region size of box b in elements: count + worst-case shift, rounded up to the alignment unit
template <typename T, int L>
__host__ __device__ constexpr int tma_box_region_size(int b) {
    constexpr int AL = 16 / (int)sizeof(T);
    return (tma_box_count<L>(b) + AL - 1 + AL - 1) / AL * AL;
}
template <typename T, int L>
__host__ __device__ constexpr int tma_box_region(int b) {  // start of region b; b = 27 gives the buffer size
    int r = 0;
    for (int j = 0; j < b; ++j) r += tma_box_region_size<T, L>(j);
    return r;
}

// ---- PTX wrappers ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n .reg .pred p;\n WAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra DONE_%=;\n bra WAIT_%=;\n DONE_%=:\n}\n" ::"r"(
            smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void bulk_s2g(void *gmem_dst, const void *smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n" ::"l"(gmem_dst), "r"(smem_u32(smem_src)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_red_add(double *gmem_dst, const double *smem_src, uint32_t bytes) {
    asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f64 [%0], [%1], %2;\n" ::"l"(gmem_dst), "r"(smem_u32(smem_src)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_red_add(float *gmem_dst, const float *smem_src, uint32_t bytes) {
    asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;\n" ::"l"(gmem_dst), "r"(smem_u32(smem_src)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

template <typename T>
__device__ __forceinline__ T lds_shared(uint32_t addr);
template <>
__device__ __forceinline__ double lds_shared<double>(uint32_t addr) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];\n" : "=d"(v) : "r"(addr));
    return v;
}
template <>
__device__ __forceinline__ float lds_shared<float>(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];\n" : "=f"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts_shared(uint32_t addr, double v) { asm volatile("st.shared.f64 [%0], %1;\n" ::"r"(addr), "d"(v) : "memory"); }
__device__ __forceinline__ void sts_shared(uint32_t addr, float v) { asm volatile("st.shared.f32 [%0], %1;\n" ::"r"(addr), "f"(v) : "memory"); }

template <int Z0, int Z1, class F>
__device__ __forceinline__ void static_for(F &&f) {
    if constexpr (Z0 < Z1) {
        f(std::integral_constant<int, Z0>{});
        static_for<Z0 + 1, Z1>(f);
    }
}

template <typename T, int L, int NWARP>
struct Sshex8TmaCfg {
    static constexpr int N = L + 1;
    static constexpr int AL = 16 / (int)sizeof(T);
    static constexpr int EPW = 32 / N;                        // elements per warp-group
    // elements per element buffer, padded to a multiple of 128 bytes: consecutive slots then start on the same bank
    // and the 8-byte shift of the interior blocks (which alternates from element to element for the compact
    // numbering) spreads the rows of two slots over disjoint bank pairs
    static constexpr int BUF = (tma_box_region<T, L>(27) * (int)sizeof(T) + 127) / 128 * 128 / (int)sizeof(T);
    static constexpr size_t buf_bytes = (size_t)NWARP * EPW * BUF * sizeof(T);
    static constexpr size_t hdr_bytes = (size_t)NWARP * 2 * EPW * sizeof(Sshex8TmaHeader);
    static constexpr size_t bar_bytes = (size_t)NWARP * 4 * sizeof(uint64_t);
    static constexpr size_t smem_bytes = buf_bytes + hdr_bytes + bar_bytes;
    static constexpr int threads = NWARP * 32;
    static_assert((BUF * sizeof(T)) % 16 == 0, "element buffers must stay 16-byte aligned");
};

// boxes moved by bulk copies when edges and corners travel per node: the interior block and the 6 faces
__host__ __device__ constexpr int tma_bulk_box(int i) { return i == 0 ? 13 : (i == 1 ? 4 : (i == 2 ? 10 : (i == 3 ? 12 : (i == 4 ? 14 : (i == 5 ? 16 : 22))))); }
// box index of edge ei (0..11): free axis ei / 4, the two fixed coordinates low / high by the bits of ei % 4
__host__ __device__ constexpr int tma_edge_box(int ei) {
    const int a = ei / 4, m = ei % 4, cu = (m & 1) * 2, cv = (m >> 1) * 2;
    return a == 0 ? 1 + 3 * cu + 9 * cv : (a == 1 ? cu + 3 + 9 * cv : cu + 3 * cv + 9);
}
__host__ __device__ constexpr int tma_corner_box(int ci) { return (ci & 1) * 2 + 3 * (((ci >> 1) & 1) * 2) + 9 * (((ci >> 2) & 1) * 2); }

// PLAIN_INTERIOR: interior nodes are exclusive to their element and y is being overwritten -> plain bulk store.
// NODE_EDGES: the 12 edges and 8 corners (12 (L-1) + 8 of the (L+1)^3 nodes) travel per node (LDG / RED by the lanes)
// instead of as 20 tiny bulk copies: a bulk copy is issued by one thread at a time (~10 issue slots each).
template <typename T, int L, int NWARP, bool PLAIN_INTERIOR, bool NODE_EDGES>
__global__ void __launch_bounds__(NWARP * 32, 1)
    sshex8_laplacian_tma_kernel(const ptrdiff_t ne, const ptrdiff_t nn, const Sshex8TmaHeader *__restrict__ headers,
                                const T *__restrict__ u, T *__restrict__ y) {
    using Cfg = Sshex8TmaCfg<T, L, NWARP>;
    constexpr int N = Cfg::N, AL = Cfg::AL, EPW = Cfg::EPW, BUF = Cfg::BUF;
    constexpr int HSZ = (int)sizeof(Sshex8TmaHeader);

    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    T *bufs = reinterpret_cast<T *>(smem_raw) + (size_t)warp * EPW * BUF;
    Sshex8TmaHeader *hdrs = reinterpret_cast<Sshex8TmaHeader *>(smem_raw + Cfg::buf_bytes) + (size_t)warp * 2 * EPW;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + Cfg::buf_bytes + Cfg::hdr_bytes) + warp * 4;
    uint64_t *bar_data = bars, *bar_hdr = bars + 1;  // bar_hdr[0..1]

    const ptrdiff_t ngroups = (ne + EPW - 1) / EPW;
    const ptrdiff_t nwt = (ptrdiff_t)gridDim.x * NWARP;
    const ptrdiff_t g0 = (ptrdiff_t)blockIdx.x * NWARP + warp;
    if (g0 >= ngroups) return;
    const int rounds = (int)((ngroups - g0 + nwt - 1) / nwt);

    if (lane == 0) {
        mbar_init(bar_data, 1);
        mbar_init(bar_hdr, 1);
        mbar_init(bar_hdr + 1, 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    fence_proxy_async_smem();
    __syncwarp();

    auto group_of = [&](int k) -> ptrdiff_t { return g0 + (ptrdiff_t)k * nwt; };
    auto nvalid_of = [&](int k) -> int {
        const ptrdiff_t left = ne - group_of(k) * EPW;
        return (int)(left < (ptrdiff_t)EPW ? left : (ptrdiff_t)EPW);
    };
    auto issue_headers = [&](int k) {
        if (k >= rounds || lane != 0) return;
        const uint32_t bytes = (uint32_t)nvalid_of(k) * HSZ;
        mbar_arrive_expect_tx(bar_hdr + (k & 1), bytes);
        bulk_g2s(hdrs + (k & 1) * EPW, headers + group_of(k) * EPW, bytes, bar_hdr + (k & 1));
    };
    issue_headers(0);

    // lane b < NBULK moves one box of every slot; lanes < EPW * N do the arithmetic of row (slot, yrow)
    constexpr int NBULK = NODE_EDGES ? 7 : 27;
    const bool box_lane = lane < NBULK;
    const int box = box_lane ? (NODE_EDGES ? tma_bulk_box(lane) : lane) : 0;
    const int cnt = tma_box_count<L>(box);
    const int reg = tma_box_region<T, L>(box);
    const bool compute_lane = lane < EPW * N;
    const int slot = compute_lane ? lane / N : 0;
    const int yrow = compute_lane ? lane - slot * N : 0;
    const int nn_al = (int)(nn & ~(ptrdiff_t)(AL - 1));  // bulk transfers never touch [nn_al, ...); ids are int32 (idx_t)

    // rows this lane reads (clamped) and their y class
    int ry[3], rcy[3];
#pragma unroll
    for (int iy = 0; iy < 3; ++iy) {
        int yy = yrow + iy - 1;
        yy = yy < 0 ? 0 : (yy > L ? L : yy);
        ry[iy] = yy;
        rcy[iy] = (yy == 0) ? 0 : (yy == L ? 2 : 1);
    }

    // per-node items (NODE_EDGES): item q = j * 32 + lane -> (slot, edge/corner box, position along the edge)
    constexpr int NS = 12 * (L - 1) + 8;                       // edge + corner nodes of an element
    constexpr int NIT = NODE_EDGES ? (EPW * NS + 31) / 32 : 1;  // items per lane and group
    // packed per item: slot (4 bits, >= EPW: no item) | box << 4 (5 bits) | t << 12 (4 bits) | region start << 16
    unsigned it[NIT];
#pragma unroll
    for (int j = 0; j < NIT; ++j) {
        const int q = j * 32 + lane;
        const int sl = q / NS, r = q - sl * NS;
        int bx, t;
        if (r < 12 * (L - 1)) {
            const int ei = r / (L - 1);
            bx = tma_edge_box(ei);
            t = 1 + r - ei * (L - 1);
        } else {
            bx = tma_corner_box(r - 12 * (L - 1));
            t = 0;
        }
        it[j] = (unsigned)(sl < EPW ? sl : 15) | ((unsigned)bx << 4) | ((unsigned)t << 12) | ((unsigned)tma_box_region<T, L>(bx) << 16);
    }
    // (buffer offset, global id) of an item: edge / corner boxes have a single non-zero stride, so the offset is
    // A + (sx + sy + sz) t; the id follows from the offset because buffer and vector agree up to the region start
    auto item_slot = [&](int j) -> int { return (int)(it[j] & 15u); };
    auto item_addr = [&](const Sshex8TmaHeader *h, int j, int &off, int &id) {
        const Sshex8TmaHeader *hs = h + item_slot(j);
        const int bx = (int)((it[j] >> 4) & 31u), t = (int)((it[j] >> 12) & 15u), R = (int)(it[j] >> 16);
        const int2 w = *reinterpret_cast<const int2 *>(&hs->desc[bx][0]);
        const int A = (int)(short)(w.x & 0xffff);
        const int st = (w.x >> 16) + (int)(short)(w.y & 0xffff) + (w.y >> 16);
        const int gm = hs->gmin[bx];
        off = A + st * t;
        id = (gm & ~(AL - 1)) + (off - R);
    };

    Sshex8LaplacianRow<T, L> row;

#pragma unroll 1
    for (int k = 0; k < rounds; ++k) {
        const Sshex8TmaHeader *h = hdrs + (k & 1) * EPW;
        mbar_wait(bar_hdr + (k & 1), (uint32_t)((k >> 1) & 1));
        // the bulk stores of the previous group must have finished reading the buffers
        bulk_wait_read_all();
        __syncwarp();
        issue_headers(k + 1);
        const int nv = nvalid_of(k);

        // ---- bulk loads of this group
        // (the slot loops of the load and flush phases stay rolled: they are latency-, not issue-bound, and the
        // unrolled arithmetic below already fills most of the instruction cache)
        uint32_t mybytes = 0;
#pragma unroll 1
        for (int s = 0; s < nv; ++s) {
            if (box_lane) {
                const int gmin = h[s].gmin[box];
                const int shift = gmin & (AL - 1);
                const int start = gmin - shift;
                int len = (shift + cnt + AL - 1) & ~(AL - 1);
                if (start + len > nn_al) len = nn_al > start ? nn_al - start : 0;  // the ragged end of the vector
                mybytes += (uint32_t)len * (uint32_t)sizeof(T);
            }
        }
        const uint32_t total = __reduce_add_sync(0xffffffffu, mybytes);
        if (lane == 0) mbar_arrive_expect_tx(bar_data, total);
        __syncwarp();
#pragma unroll 1
        for (int s = 0; s < nv; ++s) {
            if (box_lane) {
                const int gmin = h[s].gmin[box];
                const int shift = gmin & (AL - 1);
                const int start = gmin - shift;
                int len = (shift + cnt + AL - 1) & ~(AL - 1);
                T *dst = bufs + s * BUF + reg;
                if (start + len > nn_al) {
                    len = nn_al > start ? nn_al - start : 0;
                    for (int id = (start + len > gmin ? start + len : gmin); id < gmin + cnt; ++id) dst[id - start] = u[id];
                }
                if (len > 0) bulk_g2s(dst, u + start, (uint32_t)len * (uint32_t)sizeof(T), bar_data);
            }
        }
        if (NODE_EDGES) {
            T val[NIT];
            int off[NIT];
#pragma unroll
            for (int j = 0; j < NIT; ++j) {
                int id = 0;
                off[j] = 0;
                val[j] = T(0);
                if (item_slot(j) < nv) {
                    item_addr(h, j, off[j], id);
                    val[j] = u[id];
                }
            }
#pragma unroll
            for (int j = 0; j < NIT; ++j)
                if (item_slot(j) < nv) bufs[item_slot(j) * BUF + off[j]] = val[j];
        }
        mbar_wait(bar_data, (uint32_t)(k & 1));
        __syncwarp();

        // ---- arithmetic: L+1 z-steps, fully unrolled (static accumulator phases, see the core), in place
        {
            const Sshex8TmaHeader *hs = h + slot;
            const uint32_t bufa = smem_u32(bufs + slot * BUF);
            const bool active = compute_lane && slot < nv;  // headers / buffers of missing slots hold garbage
            if (active) {
                T f[6];
#pragma unroll
                for (int q = 0; q < 6; ++q) f[q] = (T)hs->fff[q];
                row.init(f, yrow);
            }
            // shared-window byte addresses of the three input rows at the current z (x = 0 / x in (0, L) / x = L),
            // their z strides and the x stride of the middle piece; the same for the output row, one plane behind
            uint32_t ra[3][3], oa[3];
            int rsz[3][3], rs1[3], osz[3], os1;
            auto load_desc = [&](int cz, int cy, int cx, int yy, int z, uint32_t &a, int &sz, int &sx) {
                const int2 w = *reinterpret_cast<const int2 *>(&hs->desc[cz * 9 + cy * 3 + cx][0]);
                const int A = (int)(short)(w.x & 0xffff), dsx = w.x >> 16;
                const int dsy = (int)(short)(w.y & 0xffff), dsz = w.y >> 16;
                a = bufa + (uint32_t)((A + dsx * (cx == 2 ? L : 0) + dsy * yy + dsz * z) * (int)sizeof(T));
                sz = dsz * (int)sizeof(T);
                sx = dsx * (int)sizeof(T);
            };
            auto store_row = [&](const T(&fin)[N]) {
                sts_shared(oa[0], fin[0]);
#pragma unroll
                for (int x = 1; x < L; ++x) sts_shared(oa[1] + (uint32_t)(os1 * x), fin[x]);
                sts_shared(oa[2], fin[L]);
            };
            static_for<0, N>([&](auto zc) {
                constexpr int z = decltype(zc)::value;
                if constexpr (z <= 1 || z == L) {
                    constexpr int cz = (z == 0) ? 0 : (z == L ? 2 : 1);
#pragma unroll
                    for (int iy = 0; iy < 3; ++iy) {
                        int sxu;
                        load_desc(cz, rcy[iy], 0, ry[iy], z, ra[iy][0], rsz[iy][0], sxu);
                        load_desc(cz, rcy[iy], 1, ry[iy], z, ra[iy][1], rsz[iy][1], rs1[iy]);
                        load_desc(cz, rcy[iy], 2, ry[iy], z, ra[iy][2], rsz[iy][2], sxu);
                    }
                } else {
#pragma unroll
                    for (int iy = 0; iy < 3; ++iy)
#pragma unroll
                        for (int c = 0; c < 3; ++c) ra[iy][c] += (uint32_t)rsz[iy][c];
                }
                if (active) {
                    row.template step_static<z>([&](const int iy, T(&r)[N]) {
                        r[0] = lds_shared<T>(ra[iy][0]);
#pragma unroll
                        for (int x = 1; x < L; ++x) r[x] = lds_shared<T>(ra[iy][1] + (uint32_t)(rs1[iy] * x));
                        r[L] = lds_shared<T>(ra[iy][2]);
                    });
                }
                __syncwarp();  // every lane has consumed plane z before anybody overwrites it (next step)
                if constexpr (z > 0) {
                    constexpr int zo = z - 1;
                    if constexpr (zo <= 1) {
                        int sxu;
                        load_desc(zo, rcy[1], 0, yrow, zo, oa[0], osz[0], sxu);
                        load_desc(zo, rcy[1], 1, yrow, zo, oa[1], osz[1], os1);
                        load_desc(zo, rcy[1], 2, yrow, zo, oa[2], osz[2], sxu);
                    } else {
#pragma unroll
                        for (int c = 0; c < 3; ++c) oa[c] += (uint32_t)osz[c];
                    }
                    if (active) {
                        store_row(row.template finished_plane<z>());
                        row.template recycle<z>();
                    }
                }
            });
            {  // output plane L: the "current" plane of the last step
                int sxu;
                load_desc(2, rcy[1], 0, yrow, L, oa[0], osz[0], sxu);
                load_desc(2, rcy[1], 1, yrow, L, oa[1], osz[1], os1);
                load_desc(2, rcy[1], 2, yrow, L, oa[2], osz[2], sxu);
                if (active) store_row(row.template finished_plane<L + 1>());
            }
        }
        __syncwarp();

        // ---- flush: zero the ride-along slots, then one bulk store / reduce per box
#pragma unroll 1
        for (int s = 0; s < nv; ++s) {
            if (box_lane) {
                const int shift = h[s].gmin[box] & (AL - 1);
                const int len = (shift + cnt + AL - 1) & ~(AL - 1);
                T *src = bufs + s * BUF + reg;
                if (!(PLAIN_INTERIOR && box == 13)) {
                    for (int j = 0; j < shift; ++j) src[j] = T(0);
                    for (int j = shift + cnt; j < len; ++j) src[j] = T(0);
                }
            }
        }
        fence_proxy_async_smem();
        __syncwarp();
#pragma unroll 1
        for (int s = 0; s < nv; ++s) {
            if (box_lane) {
                const int gmin = h[s].gmin[box];
                const int shift = gmin & (AL - 1);
                const int start = gmin - shift;
                T *src = bufs + s * BUF + reg;
                if (PLAIN_INTERIOR && box == 13) {
                    const int a0 = (gmin + AL - 1) & ~(AL - 1);
                    const int a1 = (gmin + cnt) & ~(AL - 1);
                    if (a1 > a0) bulk_s2g(y + a0, src + (a0 - start), (uint32_t)(a1 - a0) * (uint32_t)sizeof(T));
                    for (int id = gmin; id < a0 && id < gmin + cnt; ++id) y[id] = src[id - start];
                    for (int id = (a1 > a0 ? a1 : a0); id < gmin + cnt; ++id) y[id] = src[id - start];
                } else {
                    int len = (shift + cnt + AL - 1) & ~(AL - 1);
                    if (start + len > nn_al) {
                        len = nn_al > start ? nn_al - start : 0;
                        for (int id = (start + len > gmin ? start + len : gmin); id < gmin + cnt; ++id) red_add(&y[id], src[id - start]);
                    }
                    if (len > 0) bulk_red_add(y + start, src, (uint32_t)len * (uint32_t)sizeof(T));
                }
            }
        }
        bulk_commit();
        if (NODE_EDGES) {
#pragma unroll
            for (int j = 0; j < NIT; ++j) {
                if (item_slot(j) < nv) {
                    int off, id;
                    item_addr(h, j, off, id);
                    red_add(&y[id], bufs[item_slot(j) * BUF + off]);
                }
            }
        }
    }
    bulk_wait_all();
}

}  // namespace b200fem
