// Atomic-free scatter for the thread-per-element operators (tet4 / hex8): element PATCHES staged in shared memory.
//
// The per-element kernels gather NXE values and scatter NXE results with one 8-byte access each: every access moves
// a 32-byte L2 sector and every result is a RED.ADD, so the L2 / atomic units limit them at 13-26 % of the HBM
// roofline although the streams are perfectly coalesced. A patch is a run of E consecutive elements. At plan creation
// we record, per patch: the sorted list of the U unique nodes it touches, the 16-bit LOCAL index of every element
// corner, and a node -> (corner, element) incidence list in CSR form. The kernel then
//   1. stages x[unique nodes] in shared memory (U coalesced-ish loads instead of NXE * E gathers),
//   2. runs the element operator from shared memory and writes the NXE * BS results of every element to its own
//      shared-memory slots (no conflict: slot = (component, corner, element)),
//   3. sums, per unique node, the slots of its incidence list in a fixed order and issues ONE RED.ADD per
//      (node, component) -- 4 E / U ~ 12 times fewer global atomics for tet4; nodes interior to a patch could use
//      plain stores, the REDs that remain are the patch-boundary nodes.
// The sum order inside a patch is fixed, so two applies differ only in the order of the cross-patch REDs.
// Index traffic: 2 * NXE * 2 B per element (local indices + incidence slots) + 6 B per unique node, against
// NXE * 4 B of global indices -- the same bytes, all of them coalesced streams.
//
// Locality is the caller's: patches follow the element order. The reference reorders its meshes along a space-filling
// curve before use (smesh::SFC::reorder); sfem_b200/mesh.py offers the same for the synthetic boxes.
#pragma once
#include <algorithm>
#include <atomic>
#include <thread>
#include <vector>

#include "common.cuh"

namespace b200fem {

struct PatchArrays {
    ptrdiff_t ne, nchunks;
    const uint16_t *lidx;      // [nchunks][NXE][E] local node index of every element corner
    const int32_t *unodes;     // unique node ids: patch c owns the block [chunk_uoff[c], chunk_uoff[c + 1]) of
                               // roundup8(U_c + 1) entries (16-byte aligned for cp.async), the first U_c in use
    const int64_t *chunk_uoff; // [nchunks + 1] block offsets (multiples of 8), shared by unodes and noff
    const int32_t *chunk_u;    // [nchunks] U_c
    const uint16_t *nslot;     // [nchunks][NXE * E] incidence slots (corner * E + local element), grouped by node
    const uint16_t *noff;      // per patch U_c + 1 offsets into its nslot block, in the same blocks as unodes
};

struct PatchPlan {
    int nxe = 0, E = 0, max_u = 0;
    ptrdiff_t ne = 0, nchunks = 0;
    double avg_u = 0;
    DeviceArray<uint16_t> lidx, nslot, noff;
    DeviceArray<int32_t> unodes, chunk_u;
    DeviceArray<int64_t> chunk_uoff;
    bool ready = false;

    PatchArrays arrays() const {
        return PatchArrays{ne, nchunks, lidx.ptr, unodes.ptr, chunk_uoff.ptr, chunk_u.ptr, nslot.ptr, noff.ptr};
    }
    int max_u_pad() const { return (max_u + 1 + 7) / 8 * 8; }

    // elements: host SoA [nxe][ne]
    int build(int nxe_, int E_, ptrdiff_t ne_, const int32_t *const *elements) {
        nxe = nxe_;
        E = E_;
        ne = ne_;
        nchunks = (ne + E - 1) / E;
        if ((size_t)nxe * E > 65535) B2_FAIL("patch plan: NXE * E must fit 16 bits");
        std::vector<uint16_t> h_lidx((size_t)nchunks * nxe * E, 0), h_nslot((size_t)nchunks * nxe * E, 0), h_noff;
        std::vector<int32_t> h_unodes;
        std::vector<int64_t> h_uoff((size_t)nchunks + 1, 0);
        // patches are independent: build them on all host threads (a 7.4e7-tet plan takes ~2 s instead of ~15 s),
        // each thread keeping the variable-length parts (unique nodes, CSR offsets) of its patches until the prefix sum
        std::vector<std::vector<int32_t>> c_nodes((size_t)nchunks);
        std::vector<std::vector<uint16_t>> c_noff((size_t)nchunks);
        std::atomic<bool> too_many{false};
        auto work = [&](ptrdiff_t c_begin, ptrdiff_t c_end) {
            std::vector<int32_t> cnt, fill;
            for (ptrdiff_t c = c_begin; c < c_end; ++c) {
                const ptrdiff_t e0 = c * E, nel = std::min<ptrdiff_t>(E, ne - e0);
                std::vector<int32_t> &nodes = c_nodes[(size_t)c];
                for (int v = 0; v < nxe; ++v) nodes.insert(nodes.end(), elements[v] + e0, elements[v] + e0 + nel);
                std::sort(nodes.begin(), nodes.end());
                nodes.erase(std::unique(nodes.begin(), nodes.end()), nodes.end());
                const int U = (int)nodes.size();
                if (U > 65535) {
                    too_many = true;
                    return;
                }
                cnt.assign((size_t)U + 1, 0);
                uint16_t *li = &h_lidx[(size_t)c * nxe * E];
                for (int v = 0; v < nxe; ++v)
                    for (ptrdiff_t e = 0; e < nel; ++e) {
                        const int j = (int)(std::lower_bound(nodes.begin(), nodes.end(), elements[v][e0 + e]) - nodes.begin());
                        li[(size_t)v * E + e] = (uint16_t)j;
                        cnt[(size_t)j + 1]++;
                    }
                for (int j = 0; j < U; ++j) cnt[(size_t)j + 1] += cnt[(size_t)j];
                c_noff[(size_t)c].resize((size_t)U + 1);
                for (int j = 0; j <= U; ++j) c_noff[(size_t)c][(size_t)j] = (uint16_t)cnt[(size_t)j];
                uint16_t *ns = &h_nslot[(size_t)c * nxe * E];
                fill.assign(cnt.begin(), cnt.end() - 1);
                for (int v = 0; v < nxe; ++v)
                    for (ptrdiff_t e = 0; e < nel; ++e) {
                        const int j = li[(size_t)v * E + e];
                        ns[(size_t)fill[(size_t)j]++] = (uint16_t)(v * E + e);
                    }
            }
        };
        {
            const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
            const ptrdiff_t nthreads = std::max<ptrdiff_t>(1, std::min<ptrdiff_t>((ptrdiff_t)hw, nchunks / 64));
            std::vector<std::thread> pool;
            for (ptrdiff_t t = 1; t < nthreads; ++t) pool.emplace_back(work, nchunks * t / nthreads, nchunks * (t + 1) / nthreads);
            work(0, nchunks / nthreads);
            for (auto &th : pool) th.join();
        }
        if (too_many) B2_FAIL("patch plan: more than 65535 unique nodes in one patch");
        max_u = 0;
        std::vector<int32_t> h_u((size_t)nchunks, 0);
        size_t total_u = 0;
        for (ptrdiff_t c = 0; c < nchunks; ++c) {
            const int U = (int)c_nodes[(size_t)c].size();
            max_u = std::max(max_u, U);
            h_u[(size_t)c] = U;
            total_u += (size_t)U;
            h_uoff[(size_t)c + 1] = h_uoff[(size_t)c] + (U + 1 + 7) / 8 * 8;
        }
        h_unodes.assign((size_t)h_uoff[(size_t)nchunks], 0);
        h_noff.assign((size_t)h_uoff[(size_t)nchunks], 0);
        for (ptrdiff_t c = 0; c < nchunks; ++c) {
            std::copy(c_nodes[(size_t)c].begin(), c_nodes[(size_t)c].end(), h_unodes.begin() + h_uoff[(size_t)c]);
            std::copy(c_noff[(size_t)c].begin(), c_noff[(size_t)c].end(), h_noff.begin() + h_uoff[(size_t)c]);
        }
        avg_u = nchunks ? double(total_u) / double(nchunks) : 0;
        if (lidx.upload(h_lidx.data(), h_lidx.size()) || nslot.upload(h_nslot.data(), h_nslot.size()) ||
            noff.upload(h_noff.data(), h_noff.size()) || unodes.upload(h_unodes.data(), h_unodes.size()) ||
            chunk_u.upload(h_u.data(), h_u.size()) || chunk_uoff.upload(h_uoff.data(), h_uoff.size()))
            return B200FEM_FAILURE;
        ready = true;
        return B200FEM_SUCCESS;
    }

    // streamed index bytes per apply
    double index_bytes() const {
        return double(lidx.bytes() + nslot.bytes() + noff.bytes() + unodes.bytes() + chunk_uoff.bytes() + chunk_u.bytes());
    }
    // one index stage of the pipelined kernel: unodes (4 B) + noff (2 B) per padded node, lidx + nslot (2 B each) per corner
    size_t idx_stage_bytes() const { return (size_t)max_u_pad() * 6 + (size_t)nxe * E * 4; }
    // 3 index stages + 2 value stages + the per-element result slots
    size_t smem_bytes(int bs, int real_size) const {
        return 3 * idx_stage_bytes() + ((size_t)2 * max_u_pad() * bs + (size_t)nxe * bs * E) * real_size;
    }
    size_t smem_bytes_direct(int bs, int real_size) const { return ((size_t)max_u * bs + (size_t)nxe * bs * E) * real_size; }
};

// Local: __device__ void operator()(ptrdiff_t e, const T (&uin)[NXE * BS], T (&out)[NXE * BS]) const
// with uin / out ordered [corner][component]. EPT = elements per thread per patch = E / BT.
// Variant without the incidence lists: results are accumulated with shared-memory atomics into one slot per unique
// node (ATOMS, order not fixed), then one RED per (node, component). Streams 2 * NXE bytes per element less.
template <typename T, int NXE, int BS, int E, int BT, class Local>
__global__ void __launch_bounds__(BT)
    patch_apply_smem_atomic_kernel(const PatchArrays p, const Local local, const int max_u, const T *__restrict__ u,
                                   T *__restrict__ y) {
    extern __shared__ __align__(16) unsigned char patch_smem[];
    T *xs = reinterpret_cast<T *>(patch_smem);  // [max_u][BS]
    T *ys = xs + (size_t)max_u * BS;             // [max_u][BS]
    const int tid = threadIdx.x;
    for (ptrdiff_t c = blockIdx.x; c < p.nchunks; c += gridDim.x) {
        const int64_t u0 = p.chunk_uoff[c];
        const int U = p.chunk_u[c];
        const ptrdiff_t e0 = c * E;
        const int nel = (int)((p.ne - e0) < (ptrdiff_t)E ? (p.ne - e0) : (ptrdiff_t)E);
        for (int j = tid; j < U; j += BT) {
            const ptrdiff_t node = p.unodes[u0 + j];
#pragma unroll
            for (int a = 0; a < BS; ++a) {
                xs[j * BS + a] = u[node * BS + a];
                ys[j * BS + a] = T(0);
            }
        }
        __syncthreads();
        const uint16_t *li = p.lidx + (size_t)c * NXE * E;
        for (int e = tid; e < nel; e += BT) {
            T uin[NXE * BS], out[NXE * BS];
            int jj[NXE];
#pragma unroll
            for (int v = 0; v < NXE; ++v) {
                jj[v] = ldg_stream(&li[v * E + e]);
#pragma unroll
                for (int a = 0; a < BS; ++a) uin[v * BS + a] = xs[jj[v] * BS + a];
            }
            typename Local::Geo geo;
            local.load(e0 + e, geo);
            local.compute(geo, uin, out);
#pragma unroll
            for (int v = 0; v < NXE; ++v)
#pragma unroll
                for (int a = 0; a < BS; ++a) atomicAdd(&ys[jj[v] * BS + a], out[v * BS + a]);
        }
        __syncthreads();
        for (int j = tid; j < U; j += BT) {
            const ptrdiff_t node = p.unodes[u0 + j];
#pragma unroll
            for (int a = 0; a < BS; ++a) red_add(&y[node * BS + a], ys[j * BS + a]);
        }
        __syncthreads();
    }
}

__device__ __forceinline__ void patch_cp16(void *smem, const void *gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
template <typename T>
__device__ __forceinline__ void patch_cp_value(T *smem, const T *gmem) {
    if (sizeof(T) == 8)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
    else
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}

// The direct form: index lists read from global memory inside the phases, 36 KB (tet4) of shared memory per CTA. Lower
// latency tolerance per CTA, but 5 CTAs per SM: for the light operators (tet4 / hex8 Laplacian, ~45 - 90 flop per element,
// the FFF stream is 60 % of the bytes and is read by the element phase itself) the extra CTAs keep more bytes in flight
// than the staged kernel below does with its 2 CTAs, and it measures faster there (tet4 260^3: 1.56 against 2.07 ms).
// STAGE_LISTS: the incidence lists of phase 3 (slots + CSR offsets) are copied into shared memory by cp.async at the start
// of the patch and land behind phases 1 and 2, so the per-node walk of phase 3 reads shared memory instead of a chain of
// dependent global loads (24 per node for tet4).
template <typename T, int NXE, int BS, int E, int BT, class Local, bool STAGE_LISTS = false, bool HOIST_LOADS = false>
__global__ void __launch_bounds__(BT)
    patch_apply_direct_kernel(const PatchArrays p, const Local local, const int max_u, const T *__restrict__ u, T *__restrict__ y) {
    extern __shared__ __align__(16) unsigned char patch_smem[];
    constexpr int NC = NXE * E;
    uint16_t *s_ns = reinterpret_cast<uint16_t *>(patch_smem);  // [NC] + [up] when STAGE_LISTS
    const int up = (max_u + 1 + 7) / 8 * 8;
    uint16_t *s_no = s_ns + NC;
    T *xs = reinterpret_cast<T *>(patch_smem + (STAGE_LISTS ? ((size_t)NC + up) * 2 : 0));   // [max_u][BS]
    T *contrib = xs + (size_t)max_u * BS;         // [BS][NXE][E]
    const int tid = threadIdx.x;
    for (ptrdiff_t c = blockIdx.x; c < p.nchunks; c += gridDim.x) {
        const int64_t u0 = p.chunk_uoff[c];
        const int U = p.chunk_u[c];
        const ptrdiff_t e0 = c * E;
        const int nel = (int)((p.ne - e0) < (ptrdiff_t)E ? (p.ne - e0) : (ptrdiff_t)E);
        if (STAGE_LISTS) {
            const int4 *g_ns = reinterpret_cast<const int4 *>(p.nslot + (size_t)c * NC);
            const int4 *g_no = reinterpret_cast<const int4 *>(p.noff + u0);
            for (int i = tid; i < NC * 2 / 16; i += BT) patch_cp16(reinterpret_cast<int4 *>(s_ns) + i, g_ns + i);
            for (int i = tid; i < ((U + 1) * 2 + 15) / 16; i += BT) patch_cp16(reinterpret_cast<int4 *>(s_no) + i, g_no + i);
            asm volatile("cp.async.commit_group;\n" ::: "memory");
        }
        // 0. the element phase's own global loads (geometry + local corner indices of the E / BT elements of this thread, indices
        // clamped instead of predicated: a clamped duplicate recomputes the last element and rewrites its slots with the same
        // values). HOIST_LOADS: issued before the node gather instead of behind its barrier, so the two latencies overlap
        // (ncu of the tet4 instance: 12 long-scoreboard stall cycles per issue) at the price of holding them in registers.
        const uint16_t *li = p.lidx + (size_t)c * NXE * E;
        static_assert(E % BT == 0, "a patch is a whole number of elements per thread");
        constexpr int EPT = E / BT;
        typename Local::Geo geo[EPT];
        int jj[EPT][NXE];
        auto load_elements = [&]() {
#pragma unroll
            for (int k = 0; k < EPT; ++k) {
                const int e = min(tid + k * BT, nel - 1);
                local.load(e0 + e, geo[k]);
#pragma unroll
                for (int v = 0; v < NXE; ++v) jj[k][v] = ldg_stream(&li[v * E + e]);
            }
        };
        if (HOIST_LOADS) load_elements();
        // 1. stage the unique nodes' values
        for (int j = tid; j < U; j += BT) {
            const ptrdiff_t node = p.unodes[u0 + j];
#pragma unroll
            for (int a = 0; a < BS; ++a) xs[j * BS + a] = u[node * BS + a];
        }
        __syncthreads();
        // 2. element operator from shared memory into per-element slots
        if (!HOIST_LOADS) load_elements();
#pragma unroll
        for (int k = 0; k < EPT; ++k) {
            const int e = min(tid + k * BT, nel - 1);
            T uin[NXE * BS], out[NXE * BS];
#pragma unroll
            for (int v = 0; v < NXE; ++v) {
#pragma unroll
                for (int a = 0; a < BS; ++a) uin[v * BS + a] = xs[jj[k][v] * BS + a];
            }
            local.compute(geo[k], uin, out);
#pragma unroll
            for (int v = 0; v < NXE; ++v)
#pragma unroll
                for (int a = 0; a < BS; ++a) contrib[(a * NXE + v) * E + e] = out[v * BS + a];
        }
        if (STAGE_LISTS) asm volatile("cp.async.wait_group 0;\n" ::: "memory");
        __syncthreads();
        // 3. per unique node: fixed-order sum of its incidence slots, one RED per (node, component)
        const uint16_t *ns = STAGE_LISTS ? s_ns : p.nslot + (size_t)c * NXE * E;
        const uint16_t *no = STAGE_LISTS ? s_no : p.noff + u0;
        for (int j = tid; j < U; j += BT) {
            const int b = no[j], eend = no[j + 1];
            T acc[BS];
#pragma unroll
            for (int a = 0; a < BS; ++a) acc[a] = T(0);
            for (int s = b; s < eend; ++s) {
                const int slot = STAGE_LISTS ? (int)ns[s] : (int)ldg_stream(&ns[s]);
#pragma unroll
                for (int a = 0; a < BS; ++a) acc[a] += contrib[a * NXE * E + slot];
            }
            const ptrdiff_t node = p.unodes[u0 + j];
#pragma unroll
            for (int a = 0; a < BS; ++a) red_add(&y[node * BS + a], acc[a]);
        }
        __syncthreads();
    }
}

// Software-pipelined patch kernel. The first version of this kernel read its index lists (local corner indices, incidence
// slots, CSR offsets) straight from global memory inside the compute and summation phases: every unique node walked its
// incidence list through a chain of dependent global loads (8 for hex8, ~24 for tet4), three CTA-wide barriers per
// patch exposed that latency, and with 3 CTAs per SM nothing covered it (ncu of the hex8 elasticity patch kernel: 8.5
// long-scoreboard stall cycles per issue, L2 at 21 %, 349 us against 296 us for plain REDs). Here a persistent CTA keeps
//   * the index lists of patches k, k+1, k+2 in three shared-memory stages (16-byte cp.async of the padded blocks),
//   * the unique-node values of patches k and k+1 in two stages (8-byte cp.async gathers issued one patch ahead, from
//     the node list that is already in shared memory),
// so phases 2 and 3 of patch k touch shared memory only and every global access of patches k+1 / k+2 is in flight
// behind them. One wait + barrier per patch.
template <typename T, int NXE, int BS, int E, int BT, class Local>
__global__ void __launch_bounds__(BT)
    patch_apply_kernel(const PatchArrays p, const Local local, const int up, const T *__restrict__ u, T *__restrict__ y) {
    extern __shared__ __align__(16) unsigned char patch_smem[];
    constexpr int NC = NXE * E;                       // corners per patch
    const size_t idx_bytes = (size_t)up * 6 + (size_t)NC * 4;  // multiple of 16: up % 8 == 0, NC * 2 % 16 == 0
    static_assert((NC * 2) % 16 == 0, "corner lists must be 16-byte multiples");
    auto st_unodes = [&](int s) { return reinterpret_cast<int32_t *>(patch_smem + (size_t)s * idx_bytes); };
    auto st_noff = [&](int s) { return reinterpret_cast<uint16_t *>(patch_smem + (size_t)s * idx_bytes + (size_t)up * 4); };
    auto st_lidx = [&](int s) { return reinterpret_cast<uint16_t *>(patch_smem + (size_t)s * idx_bytes + (size_t)up * 6); };
    auto st_nslot = [&](int s) { return st_lidx(s) + NC; };
    T *xs_base = reinterpret_cast<T *>(patch_smem + 3 * idx_bytes);  // [2][up][BS]
    T *contrib = xs_base + (size_t)2 * up * BS;                       // [BS][NXE][E]
    const int tid = threadIdx.x;
    const ptrdiff_t stride = gridDim.x;

    auto issue_idx = [&](const ptrdiff_t c, const int s) {
        if (c >= p.nchunks) return;
        const int64_t u0 = p.chunk_uoff[c];
        const int U = p.chunk_u[c];
        const int4 *g_un = reinterpret_cast<const int4 *>(p.unodes + u0);
        const int4 *g_no = reinterpret_cast<const int4 *>(p.noff + u0);
        const int4 *g_li = reinterpret_cast<const int4 *>(p.lidx + (size_t)c * NC);
        const int4 *g_ns = reinterpret_cast<const int4 *>(p.nslot + (size_t)c * NC);
        int4 *s_un = reinterpret_cast<int4 *>(st_unodes(s)), *s_no = reinterpret_cast<int4 *>(st_noff(s));
        int4 *s_li = reinterpret_cast<int4 *>(st_lidx(s)), *s_ns = reinterpret_cast<int4 *>(st_nslot(s));
        for (int i = tid; i < (U * 4 + 15) / 16; i += BT) patch_cp16(s_un + i, g_un + i);
        for (int i = tid; i < ((U + 1) * 2 + 15) / 16; i += BT) patch_cp16(s_no + i, g_no + i);
        for (int i = tid; i < NC * 2 / 16; i += BT) {
            patch_cp16(s_li + i, g_li + i);
            patch_cp16(s_ns + i, g_ns + i);
        }
    };
    auto issue_x = [&](const ptrdiff_t c, const int s_idx, const int s_x) {  // node list of c is in stage s_idx
        if (c >= p.nchunks) return;
        const int U = p.chunk_u[c];
        const int32_t *un = st_unodes(s_idx);
        T *xs = xs_base + (size_t)s_x * up * BS;
        for (int j = tid; j < U; j += BT) {
            const ptrdiff_t node = un[j];
#pragma unroll
            for (int a = 0; a < BS; ++a) patch_cp_value(&xs[j * BS + a], &u[node * BS + a]);
        }
    };
    auto commit = [] { asm volatile("cp.async.commit_group;\n" ::: "memory"); };
    auto wait_all = [] { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); };

    const ptrdiff_t c0 = blockIdx.x;
    issue_idx(c0, 0);
    issue_idx(c0 + stride, 1);
    commit();
    wait_all();
    __syncthreads();
    issue_x(c0, 0, 0);
    commit();
    wait_all();
    __syncthreads();

    // E == BT (one element per thread and patch): the geometry of the NEXT patch's element is loaded into registers while the
    // current one computes -- it is the one global load of the element phase that the staging above does not cover (ncu of the
    // hex8 elasticity instance: 2.4 long-scoreboard stall cycles per issue with 4 warps per scheduler)
    constexpr bool kPrefetchGeo = (E == BT);
    typename Local::Geo geo_cur, geo_next;
    if (kPrefetchGeo && c0 < p.nchunks) {
        const ptrdiff_t last = p.ne - 1, e = c0 * E + tid;
        local.load(e < last ? e : last, geo_cur);
    }
    int k = 0;
    for (ptrdiff_t c = c0; c < p.nchunks; c += stride, ++k) {
        const int si = k % 3, sx = k & 1;
        issue_idx(c + 2 * stride, (k + 2) % 3);
        issue_x(c + stride, (k + 1) % 3, (k + 1) & 1);
        commit();
        if (kPrefetchGeo && c + stride < p.nchunks) {
            const ptrdiff_t last = p.ne - 1, e = (c + stride) * E + tid;
            local.load(e < last ? e : last, geo_next);
        }

        const int U = p.chunk_u[c];
        const ptrdiff_t e0 = c * E;
        const int nel = (int)((p.ne - e0) < (ptrdiff_t)E ? (p.ne - e0) : (ptrdiff_t)E);
        const T *xs = xs_base + (size_t)sx * up * BS;
        // 2. element operator from shared memory into per-element slots
        const uint16_t *li = st_lidx(si);
        for (int e = tid; e < nel; e += BT) {
            typename Local::Geo geo;
            if (kPrefetchGeo)
                geo = geo_cur;
            else
                local.load(e0 + e, geo);
            T uin[NXE * BS], out[NXE * BS];
#pragma unroll
            for (int v = 0; v < NXE; ++v) {
                const int j = li[v * E + e];
#pragma unroll
                for (int a = 0; a < BS; ++a) uin[v * BS + a] = xs[j * BS + a];
            }
            local.compute(geo, uin, out);
#pragma unroll
            for (int v = 0; v < NXE; ++v)
#pragma unroll
                for (int a = 0; a < BS; ++a) contrib[(a * NXE + v) * E + e] = out[v * BS + a];
        }
        if (kPrefetchGeo) geo_cur = geo_next;
        __syncthreads();
        // 3. per unique node: fixed-order sum of its incidence slots, one RED per (node, component)
        const uint16_t *ns = st_nslot(si);
        const uint16_t *no = st_noff(si);
        const int32_t *un = st_unodes(si);
        for (int j = tid; j < U; j += BT) {
            const int b = no[j], eend = no[j + 1];
            T acc[BS];
#pragma unroll
            for (int a = 0; a < BS; ++a) acc[a] = T(0);
            for (int s = b; s < eend; ++s) {
                const int slot = ns[s];
#pragma unroll
                for (int a = 0; a < BS; ++a) acc[a] += contrib[a * NXE * E + slot];
            }
            const ptrdiff_t node = un[j];
#pragma unroll
            for (int a = 0; a < BS; ++a) red_add(&y[node * BS + a], acc[a]);
        }
        wait_all();  // index lists of patch k + 2, values of patch k + 1
        __syncthreads();
    }
}

template <typename T, int NXE, int BS, int E, int BT, class Local>
int launch_patch_apply_smem_atomic(const PatchPlan &plan, const Local &local, const T *x, T *y, cudaStream_t s) {
    auto kern = patch_apply_smem_atomic_kernel<T, NXE, BS, E, BT, Local>;
    const size_t smem = (size_t)plan.max_u * BS * 2 * sizeof(T);
    static SmemOptIn opt;
    if (opt.ensure(kern, smem)) return B200FEM_FAILURE;
    if (plan.nchunks == 0) return B200FEM_SUCCESS;
    const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(8, (200 * 1024) / std::max<size_t>(smem, 1)));
    const int grid = (int)std::min<ptrdiff_t>(plan.nchunks, (ptrdiff_t)sm_count() * per_sm);
    kern<<<grid, BT, smem, s>>>(plan.arrays(), local, plan.max_u, x, y);
    B2_CHECK_LAUNCH();
    return B200FEM_SUCCESS;
}

template <typename T, int NXE, int BS, int E, int BT, class Local, bool STAGE_LISTS = false, bool HOIST_LOADS = false>
int launch_patch_apply_direct(const PatchPlan &plan, const Local &local, const T *x, T *y, cudaStream_t s) {
    auto kern = patch_apply_direct_kernel<T, NXE, BS, E, BT, Local, STAGE_LISTS, HOIST_LOADS>;
    const size_t smem = ((size_t)plan.max_u * BS + (size_t)NXE * BS * E) * sizeof(T) +
                        (STAGE_LISTS ? ((size_t)NXE * E + (size_t)plan.max_u_pad()) * 2 : 0);
    static SmemOptIn opt;
    if (opt.ensure(kern, smem)) return B200FEM_FAILURE;
    if (plan.nchunks == 0) return B200FEM_SUCCESS;
    const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(8, (200 * 1024) / std::max<size_t>(smem, 1)));
    const int grid = (int)std::min<ptrdiff_t>(plan.nchunks, (ptrdiff_t)sm_count() * per_sm);
    kern<<<grid, BT, smem, s>>>(plan.arrays(), local, plan.max_u, x, y);
    B2_CHECK_LAUNCH();
    return B200FEM_SUCCESS;
}

template <typename T, int NXE, int BS, int E, int BT, class Local>
int launch_patch_apply(const PatchPlan &plan, const Local &local, const T *x, T *y, cudaStream_t s) {
    auto kern = patch_apply_kernel<T, NXE, BS, E, BT, Local>;
    const size_t smem = plan.smem_bytes(BS, (int)sizeof(T));
    static SmemOptIn opt;
    if (opt.ensure(kern, smem)) return B200FEM_FAILURE;
    if (plan.nchunks == 0) return B200FEM_SUCCESS;
    const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(2048 / BT, (220 * 1024) / (smem + 1024)));
    const int grid = (int)std::min<ptrdiff_t>(plan.nchunks, (ptrdiff_t)sm_count() * per_sm);
    kern<<<grid, BT, smem, s>>>(plan.arrays(), local, plan.max_u_pad(), x, y);
    B2_CHECK_LAUNCH();
    return B200FEM_SUCCESS;
}

}  // namespace b200fem
