// Packed meshes: the Laplacian on smesh::PackedMesh<uint16_t> connectivity.
//
// Replaces PackedLaplacian::apply -> packed_laplacian_apply<uint16_t> -> PackedLaplacianApply<pack_idx_t, NXE, MicroKernel>::apply
// (frontend/ops/sfem_PackedLaplacian.cpp:111-176, dispatch :551-612, Op :875-930) for TET4 / HEX8 / TET10 (+ MACRO_TET4).
//
// The format, as that routine consumes it. Elements are cut into PACKS of n_elements_per_pack consecutive elements. The
// nodes are numbered so that the nodes a pack OWNS are the contiguous range [owned_nodes_ptr[p], owned_nodes_ptr[p + 1]),
// the last n_shared[p] of them also touched by other packs; the other nodes a pack touches are its ghosts
// ghost_idx[ghost_ptr[p] .. ghost_ptr[p + 1]). The connectivity is 16-bit and LOCAL to the pack: l < n_owned is node
// owned_nodes_ptr[p] + l, otherwise ghost l - n_owned. Index stream: 2 bytes per element corner instead of 4.
//
// smesh (and with it PackedMesh<T>::create) is absent from the reference checkout, so the packer below is ours: it
// produces exactly the arrays PackedLaplacian::apply reads (parity of the PACKER is unpinned; parity of the APPLY is
// pinned by expanding the packed connectivity to global ids and running the reference's own *_laplacian_apply_opt).
//
// Kernels. Staged form ("packed_kernel" = 0): one CTA per pack (persistent loop); the pack's node values are staged in
// shared memory -- ONE coalesced copy of the owned range + a gather of the ghosts, instead of NXE gathers per element --,
// the element operator runs out of shared memory and accumulates with shared-memory atomics, then the exclusively owned
// nodes are written with plain read-modify-writes (no other pack touches them: the reference skips its `omp atomic` for
// the same nodes) and only the shared / ghost nodes take a RED. Direct form (the default, and the only one for packs
// beyond shared memory -- the format allows 65536 nodes): one thread per element, local indices resolved to global ids
// on the fly, RED per corner.
// Both read the 2-byte index stream; measured, the direct form is the faster one on B200 (see PackedLaplacian below).
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <numeric>
#include <thread>
#include <vector>

#include "common.cuh"
#include "hex8_local.cuh"
#include "p2_tet_local.cuh"

namespace b200fem {

// ---------------------------------------------------------------------------------------------- the packer (host)
struct PackedMeshData {
    int nxe = 0;
    ptrdiff_t ne = 0, nn = 0, n_packs = 0, n_elements_per_pack = 0, max_nodes_per_pack = 0;
    std::vector<std::vector<uint16_t>> rows;  // [nxe][ne] local indices
    std::vector<const uint16_t *> row_ptrs;
    std::vector<ptrdiff_t> owned_nodes_ptr, n_shared, ghost_ptr;
    std::vector<int32_t> ghost_idx;   // new (packed) node ids
    std::vector<int32_t> new_to_old;  // node permutation: packed id -> id in the input mesh
    // CSR of the two-pass apply (PackedLaplacianApply::apply_two_pass, frontend/ops/sfem_PackedLaplacian.cpp:178-262): row r
    // collects the ghost-buffer entries ghost_reduce_idx[ghost_reduce_ptr[r] .. ghost_reduce_ptr[r + 1]) into node ghost_reduce_dest[r]
    std::vector<ptrdiff_t> ghost_reduce_ptr, ghost_reduce_idx;
    std::vector<int32_t> ghost_reduce_dest;
};

static int build_packed_mesh(PackedMeshData &pm, int nxe, ptrdiff_t ne, ptrdiff_t nn, const int32_t *const *elements, ptrdiff_t epp) {
    if (nxe < 1 || ne < 0 || nn < 0) B2_FAIL("packed mesh: bad sizes");
    if (epp <= 0) epp = 1024;
    for (int v = 0; v < nxe; ++v)
        for (ptrdiff_t e = 0; e < ne; ++e)
            if (elements[v][e] < 0 || elements[v][e] >= nn) B2_FAIL("packed mesh: node index out of range");
    pm.nxe = nxe;
    pm.ne = ne;
    pm.nn = nn;
    pm.n_elements_per_pack = epp;
    const ptrdiff_t np = (ne + epp - 1) / epp;
    pm.n_packs = np;
    // owner = the first pack that touches a node; shared = touched by a second pack
    std::vector<int32_t> owner((size_t)nn, -1);
    std::vector<uint8_t> shared((size_t)nn, 0);
    for (ptrdiff_t p = 0; p < np; ++p) {
        const ptrdiff_t e0 = p * epp, e1 = std::min(ne, e0 + epp);
        for (int v = 0; v < nxe; ++v)
            for (ptrdiff_t e = e0; e < e1; ++e) {
                const int32_t n = elements[v][e];
                if (owner[(size_t)n] < 0)
                    owner[(size_t)n] = (int32_t)p;
                else if (owner[(size_t)n] != (int32_t)p)
                    shared[(size_t)n] = 1;
            }
    }
    // nodes no element touches are appended to the last pack's range so that the ranges still partition [0, nn)
    const int32_t orphan_pack = (int32_t)std::max<ptrdiff_t>(np - 1, 0);
    // new numbering: by (owner pack, exclusive before shared, old id)
    pm.new_to_old.resize((size_t)nn);
    std::iota(pm.new_to_old.begin(), pm.new_to_old.end(), 0);
    auto pack_of = [&](int32_t n) { return owner[(size_t)n] < 0 ? orphan_pack : owner[(size_t)n]; };
    std::stable_sort(pm.new_to_old.begin(), pm.new_to_old.end(), [&](int32_t a, int32_t b) {
        const int32_t pa = pack_of(a), pb = pack_of(b);
        if (pa != pb) return pa < pb;
        return shared[(size_t)a] < shared[(size_t)b];
    });
    std::vector<int32_t> old_to_new((size_t)nn);
    for (ptrdiff_t i = 0; i < nn; ++i) old_to_new[(size_t)pm.new_to_old[(size_t)i]] = (int32_t)i;
    pm.owned_nodes_ptr.assign((size_t)np + 1, 0);
    pm.n_shared.assign((size_t)np, 0);
    for (ptrdiff_t i = 0; i < nn; ++i) {
        const int32_t o = pm.new_to_old[(size_t)i], p = pack_of(o);
        if (np == 0) break;
        pm.owned_nodes_ptr[(size_t)p + 1]++;
        if (shared[(size_t)o]) pm.n_shared[(size_t)p]++;
    }
    for (ptrdiff_t p = 0; p < np; ++p) pm.owned_nodes_ptr[(size_t)p + 1] += pm.owned_nodes_ptr[(size_t)p];
    // ghosts + local connectivity: packs are independent once the numbering is fixed -- built on all host threads (like
    // PatchPlan::build), every pack keeping its ghost list until the prefix sum
    pm.rows.assign((size_t)nxe, std::vector<uint16_t>((size_t)ne, 0));
    std::vector<std::vector<int32_t>> pack_ghosts((size_t)np);
    std::vector<ptrdiff_t> pack_local((size_t)np, 0);
    std::atomic<bool> too_many{false};
    auto work = [&](ptrdiff_t p_begin, ptrdiff_t p_end) {
        for (ptrdiff_t p = p_begin; p < p_end; ++p) {
            const ptrdiff_t e0 = p * epp, e1 = std::min(ne, e0 + epp);
            const ptrdiff_t base = pm.owned_nodes_ptr[(size_t)p], n_owned = pm.owned_nodes_ptr[(size_t)p + 1] - base;
            std::vector<int32_t> &ghosts = pack_ghosts[(size_t)p];
            for (int v = 0; v < nxe; ++v)
                for (ptrdiff_t e = e0; e < e1; ++e) {
                    const int32_t n = old_to_new[(size_t)elements[v][e]];
                    if (n < base || n >= base + n_owned) ghosts.push_back(n);
                }
            std::sort(ghosts.begin(), ghosts.end());
            ghosts.erase(std::unique(ghosts.begin(), ghosts.end()), ghosts.end());
            const ptrdiff_t n_local = n_owned + (ptrdiff_t)ghosts.size();
            if (n_local > 65536) {
                too_many = true;
                return;
            }
            pack_local[(size_t)p] = n_local;
            for (int v = 0; v < nxe; ++v)
                for (ptrdiff_t e = e0; e < e1; ++e) {
                    const int32_t n = old_to_new[(size_t)elements[v][e]];
                    ptrdiff_t l;
                    if (n >= base && n < base + n_owned)
                        l = n - base;
                    else
                        l = n_owned + (std::lower_bound(ghosts.begin(), ghosts.end(), n) - ghosts.begin());
                    pm.rows[(size_t)v][(size_t)e] = (uint16_t)l;
                }
        }
    };
    {
        const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
        const ptrdiff_t nthreads = std::max<ptrdiff_t>(1, std::min<ptrdiff_t>((ptrdiff_t)hw, np / 64));
        std::vector<std::thread> pool;
        for (ptrdiff_t t = 1; t < nthreads; ++t) pool.emplace_back(work, np * t / nthreads, np * (t + 1) / nthreads);
        work(0, np / nthreads);
        for (auto &th : pool) th.join();
    }
    if (too_many) B2_FAIL("packed mesh: a pack touches more than 65536 nodes, use fewer elements per pack");
    pm.ghost_ptr.assign((size_t)np + 1, 0);
    pm.ghost_idx.clear();
    pm.max_nodes_per_pack = 0;
    for (ptrdiff_t p = 0; p < np; ++p) {
        pm.max_nodes_per_pack = std::max(pm.max_nodes_per_pack, pack_local[(size_t)p]);
        pm.ghost_idx.insert(pm.ghost_idx.end(), pack_ghosts[(size_t)p].begin(), pack_ghosts[(size_t)p].end());
        pm.ghost_ptr[(size_t)p + 1] = (ptrdiff_t)pm.ghost_idx.size();
    }
    pm.row_ptrs.clear();
    for (int v = 0; v < nxe; ++v) pm.row_ptrs.push_back(pm.rows[(size_t)v].data());
    // ghost entries grouped by the node they stand for
    const size_t ng = pm.ghost_idx.size();
    pm.ghost_reduce_idx.resize(ng);
    std::iota(pm.ghost_reduce_idx.begin(), pm.ghost_reduce_idx.end(), (ptrdiff_t)0);
    std::stable_sort(pm.ghost_reduce_idx.begin(), pm.ghost_reduce_idx.end(),
                     [&](ptrdiff_t a, ptrdiff_t b) { return pm.ghost_idx[(size_t)a] < pm.ghost_idx[(size_t)b]; });
    pm.ghost_reduce_ptr.assign(1, 0);
    pm.ghost_reduce_dest.clear();
    for (size_t k = 0; k < ng; ++k) {
        const int32_t n = pm.ghost_idx[(size_t)pm.ghost_reduce_idx[k]];
        if (pm.ghost_reduce_dest.empty() || pm.ghost_reduce_dest.back() != n) {
            if (!pm.ghost_reduce_dest.empty()) pm.ghost_reduce_ptr.push_back((ptrdiff_t)k);
            pm.ghost_reduce_dest.push_back(n);
        }
    }
    if (ng) pm.ghost_reduce_ptr.push_back((ptrdiff_t)ng);
    return B200FEM_SUCCESS;
}

// ---------------------------------------------------------------------------------------------- element operators
enum { kFamTet4 = 0, kFamHex8 = 1, kFamTet10 = 2, kFamMacroTet4 = 3 };

template <typename T, int FAM>
struct PackedLocal;
template <typename T>
struct PackedLocal<T, kFamTet4> {
    static constexpr int NXE = 4;
    static __device__ __forceinline__ void apply(const T (&f)[6], const T (&u)[4], T (&o)[4]) {
        o[0] = o[1] = o[2] = o[3] = T(0);
        tet4_laplacian_local_add(f, u[0], u[1], u[2], u[3], o[0], o[1], o[2], o[3]);
    }
};
template <typename T>
struct PackedLocal<T, kFamHex8> {
    static constexpr int NXE = 8;
    static __device__ __forceinline__ void apply(const T (&f)[6], const T (&u)[8], T (&o)[8]) {
        T ul[8], ol[8];
#pragma unroll
        for (int v = 0; v < 8; ++v) ul[kVtkToLex[v]] = u[v];
        hex8_laplacian_local(f, ul, ol);
#pragma unroll
        for (int v = 0; v < 8; ++v) o[v] = ol[kVtkToLex[v]];
    }
};
template <typename T>
struct PackedLocal<T, kFamTet10> {
    static constexpr int NXE = 10;
    static __device__ __forceinline__ void apply(const T (&f)[6], const T (&u)[10], T (&o)[10]) { tet10_laplacian_local(f, u, o); }
};
template <typename T>
struct PackedLocal<T, kFamMacroTet4> {
    static constexpr int NXE = 10;
    static __device__ __forceinline__ void apply(const T (&f)[6], const T (&u)[10], T (&o)[10]) { macro_tet4_laplacian_local(f, u, o); }
};

struct PackedArrays {
    ptrdiff_t ne, n_packs, epp;
    const uint16_t *lidx;              // [NXE][ne]
    const float *fff;                  // SoA [6][ne]
    const int64_t *owned_nodes_ptr;    // [n_packs + 1]
    const int32_t *n_shared;           // [n_packs]
    const int64_t *ghost_ptr;          // [n_packs + 1]
    const int32_t *ghost_idx;
};

// one CTA per pack, node values and results staged in shared memory
// TWO_PASS (the reference's atomics-free variant): every owned node takes a plain read-modify-write -- in this pass only the
// owner pack writes it -- and the ghost results go to the pack's slots of `ghost_buf`; packed_ghost_reduce_kernel adds them up.
template <typename T, int FAM, int BT, bool TWO_PASS = false>
__global__ void __launch_bounds__(BT)
    packed_laplacian_staged_kernel(const PackedArrays a, const T *__restrict__ u, T *__restrict__ y, T *__restrict__ ghost_buf) {
    using L = PackedLocal<T, FAM>;
    constexpr int NXE = L::NXE;
    extern __shared__ __align__(16) unsigned char packed_smem[];
    const int tid = threadIdx.x;
    for (ptrdiff_t p = blockIdx.x; p < a.n_packs; p += gridDim.x) {
        const ptrdiff_t base = a.owned_nodes_ptr[p];
        const int n_owned = (int)(a.owned_nodes_ptr[p + 1] - base);
        const int n_excl = n_owned - a.n_shared[p];
        const ptrdiff_t g0 = a.ghost_ptr[p];
        const int n_ghost = (int)(a.ghost_ptr[p + 1] - g0);
        const int n_local = n_owned + n_ghost;
        T *xs = reinterpret_cast<T *>(packed_smem);
        T *ys = xs + n_local;
        for (int k = tid; k < n_owned; k += BT) {
            xs[k] = u[base + k];
            ys[k] = T(0);
        }
        for (int k = tid; k < n_ghost; k += BT) {
            xs[n_owned + k] = u[a.ghost_idx[g0 + k]];
            ys[n_owned + k] = T(0);
        }
        __syncthreads();
        const ptrdiff_t e0 = p * a.epp, e1 = (e0 + a.epp < a.ne) ? e0 + a.epp : a.ne;
        for (ptrdiff_t e = e0 + tid; e < e1; e += BT) {
            int l[NXE];
#pragma unroll
            for (int v = 0; v < NXE; ++v) l[v] = ldg_stream(&a.lidx[(size_t)v * a.ne + e]);
            T f[6], ul[NXE], ol[NXE];
#pragma unroll
            for (int k = 0; k < 6; ++k) f[k] = (T)ldg_stream(&a.fff[(size_t)k * a.ne + e]);
#pragma unroll
            for (int v = 0; v < NXE; ++v) ul[v] = xs[l[v]];
            L::apply(f, ul, ol);
#pragma unroll
            for (int v = 0; v < NXE; ++v) atomicAdd(&ys[l[v]], ol[v]);
        }
        __syncthreads();
        if (TWO_PASS) {
            for (int k = tid; k < n_owned; k += BT) y[base + k] += ys[k];
            for (int k = tid; k < n_ghost; k += BT) ghost_buf[g0 + k] = ys[n_owned + k];
        } else {
            for (int k = tid; k < n_excl; k += BT) y[base + k] += ys[k];  // exclusively owned: no other pack writes these
            for (int k = n_excl + tid; k < n_owned; k += BT) red_add(&y[base + k], ys[k]);
            for (int k = tid; k < n_ghost; k += BT) red_add(&y[a.ghost_idx[g0 + k]], ys[n_owned + k]);
        }
        __syncthreads();
    }
}

// second pass: one thread per shared node, fixed-order sum of its ghost-buffer entries (no atomics anywhere in global memory)
template <typename T>
__global__ void __launch_bounds__(256)
    packed_ghost_reduce_kernel(const ptrdiff_t n_rows, const int64_t *__restrict__ ptr, const int64_t *__restrict__ idx,
                               const int32_t *__restrict__ dest, const T *__restrict__ ghost_buf, T *__restrict__ y) {
    for (ptrdiff_t r = (ptrdiff_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_rows; r += (ptrdiff_t)gridDim.x * blockDim.x) {
        T sum = T(0);
        for (int64_t j = ptr[r]; j < ptr[r + 1]; ++j) sum += ghost_buf[idx[j]];
        y[dest[r]] += sum;
    }
}

// any pack size: one thread per element, local indices resolved on the fly
template <typename T, int FAM>
__global__ void __launch_bounds__(128) packed_laplacian_direct_kernel(const PackedArrays a, const T *__restrict__ u, T *__restrict__ y) {
    using L = PackedLocal<T, FAM>;
    constexpr int NXE = L::NXE;
    for (ptrdiff_t e = (ptrdiff_t)blockIdx.x * blockDim.x + threadIdx.x; e < a.ne; e += (ptrdiff_t)gridDim.x * blockDim.x) {
        const ptrdiff_t p = e / a.epp;
        const ptrdiff_t base = a.owned_nodes_ptr[p];
        const int n_owned = (int)(a.owned_nodes_ptr[p + 1] - base);
        const ptrdiff_t g0 = a.ghost_ptr[p];
        ptrdiff_t g[NXE];
#pragma unroll
        for (int v = 0; v < NXE; ++v) {
            const int l = ldg_stream(&a.lidx[(size_t)v * a.ne + e]);
            g[v] = l < n_owned ? base + l : (ptrdiff_t)a.ghost_idx[g0 + (l - n_owned)];
        }
        T f[6], ul[NXE], ol[NXE];
#pragma unroll
        for (int k = 0; k < 6; ++k) f[k] = (T)ldg_stream(&a.fff[(size_t)k * a.ne + e]);
#pragma unroll
        for (int v = 0; v < NXE; ++v) ul[v] = u[g[v]];
        L::apply(f, ul, ol);
#pragma unroll
        for (int v = 0; v < NXE; ++v) red_add(&y[g[v]], ol[v]);
    }
}

struct PackedLaplacian final : OpBase {
    int fam = kFamTet4;
    ptrdiff_t n_packs = 0, epp = 0, max_nodes = 0;
    DeviceArray<uint16_t> lidx;
    DeviceArray<float> fff;
    DeviceArray<int64_t> owned_nodes_ptr, ghost_ptr;
    DeviceArray<int32_t> n_shared, ghost_idx;
    // 1 = direct (default), 0 = staged when a pack fits shared memory. Measured on B200 (profiles/r02p2_extras.jsonl, 2.96e6
    // elements, fp64): tet4 0.082 ms direct / 0.089 ms staged, tet10 0.153 / 0.202 ms -- the vectors of such meshes live in the
    // 126 MB L2, so the staging buys no HBM traffic and its two barriers per pack cost
    int packed_kernel = 1;
    // two-pass form (b200fem_packed_laplacian_set_two_pass + option "packed_two_pass" = 1, the reference's SFEM_PACKED_TWO_PASS)
    int two_pass = 0;
    ptrdiff_t n_reduce_rows = 0;
    DeviceArray<int64_t> reduce_ptr, reduce_idx;
    DeviceArray<int32_t> reduce_dest;
    DeviceArray<unsigned char> ghost_buf;

    PackedArrays arrays() const {
        return PackedArrays{ne, n_packs, epp, lidx.ptr, fff.ptr, owned_nodes_ptr.ptr, n_shared.ptr, ghost_ptr.ptr, ghost_idx.ptr};
    }
    template <typename T, int FAM>
    int launch(const T *x, T *y, cudaStream_t s) {
        const size_t smem = (size_t)2 * max_nodes * sizeof(T);
        if (two_pass) {
            constexpr int BT = 256;
            auto kern = packed_laplacian_staged_kernel<T, FAM, BT, true>;
            static SmemOptIn opt;
            if (opt.ensure(kern, smem)) return B200FEM_FAILURE;
            const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(8, (200 * 1024) / std::max<size_t>(smem, 1)));
            const int grid = (int)std::min<ptrdiff_t>(n_packs, (ptrdiff_t)sm_count() * per_sm);
            kern<<<grid, BT, smem, s>>>(arrays(), x, y, reinterpret_cast<T *>(ghost_buf.ptr));
            B2_CHECK_LAUNCH();
            if (n_reduce_rows) {
                const int g2 = (int)std::min<ptrdiff_t>((n_reduce_rows + 255) / 256, (ptrdiff_t)sm_count() * 16);
                packed_ghost_reduce_kernel<T><<<g2, 256, 0, s>>>(n_reduce_rows, reduce_ptr.ptr, reduce_idx.ptr, reduce_dest.ptr,
                                                                reinterpret_cast<const T *>(ghost_buf.ptr), y);
                B2_CHECK_LAUNCH();
            }
            return B200FEM_SUCCESS;
        }
        if (packed_kernel == 0 && smem <= 200 * 1024) {
            constexpr int BT = 256;
            auto kern = packed_laplacian_staged_kernel<T, FAM, BT>;
            static SmemOptIn opt;
            if (opt.ensure(kern, smem)) return B200FEM_FAILURE;
            const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(8, (200 * 1024) / std::max<size_t>(smem, 1)));
            const int grid = (int)std::min<ptrdiff_t>(n_packs, (ptrdiff_t)sm_count() * per_sm);
            kern<<<grid, BT, smem, s>>>(arrays(), x, y, (T *)nullptr);
        } else {
            const int grid = (int)std::min<ptrdiff_t>((ne + 127) / 128, (ptrdiff_t)sm_count() * 16);
            packed_laplacian_direct_kernel<T, FAM><<<grid, 128, 0, s>>>(arrays(), x, y);
        }
        B2_CHECK_LAUNCH();
        return B200FEM_SUCCESS;
    }
    template <typename T>
    int apply_t(const T *x, T *y, cudaStream_t s) {
        if (ne == 0) return B200FEM_SUCCESS;
        switch (fam) {
            case kFamTet4: return launch<T, kFamTet4>(x, y, s);
            case kFamHex8: return launch<T, kFamHex8>(x, y, s);
            case kFamTet10: return launch<T, kFamTet10>(x, y, s);
            default: return launch<T, kFamMacroTet4>(x, y, s);
        }
    }
    int apply(const void *x, void *y, bool overwrite, cudaStream_t s) override {
        if (overwrite) B2_CUDA(cudaMemsetAsync(y, 0, (size_t)nn * real_type, s));
        return real_type == B200FEM_F64 ? apply_t<double>((const double *)x, (double *)y, s) : apply_t<float>((const float *)x, (float *)y, s);
    }
    int launches(bool overwrite) const override { return 1 + (two_pass && n_reduce_rows ? 1 : 0) + (overwrite ? 1 : 0); }
    double streamed_bytes(bool overwrite) const override {
        return double(lidx.bytes() + fff.bytes() + ghost_idx.bytes() + owned_nodes_ptr.bytes() + ghost_ptr.bytes() + n_shared.bytes()) +
               double(nn) * real_type * (overwrite ? 2.0 : 3.0);
    }
    int set_option(const char *key, int value) override {
        if (!strcmp(key, "packed_kernel")) { packed_kernel = value; return B200FEM_SUCCESS; }
        if (!strcmp(key, "packed_two_pass")) {
            if (value && !ghost_buf.ptr && ghost_idx.n) B2_FAIL("packed_two_pass: call b200fem_packed_laplacian_set_two_pass first");
            two_pass = value;
            return B200FEM_SUCCESS;
        }
        return OpBase::set_option(key, value);
    }
    int get_option(const char *key, int *value) override {
        if (!strcmp(key, "packed_kernel")) { *value = packed_kernel; return B200FEM_SUCCESS; }
        if (!strcmp(key, "packed_two_pass")) { *value = two_pass; return B200FEM_SUCCESS; }
        if (!strcmp(key, "packed_staged")) { *value = (packed_kernel == 0 && (size_t)2 * max_nodes * real_type <= 200 * 1024) ? 1 : 0; return B200FEM_SUCCESS; }
        return OpBase::get_option(key, value);
    }
};

OpBase *make_packed_laplacian(int element_type, ptrdiff_t n_packs, ptrdiff_t epp, ptrdiff_t ne, ptrdiff_t nn, ptrdiff_t max_nodes,
                              const uint16_t *const *elements, const float *fff, const ptrdiff_t *owned_nodes_ptr, const ptrdiff_t *n_shared,
                              const ptrdiff_t *ghost_ptr, const int32_t *ghost_idx, int real_type) {
    int fam, nxe;
    switch (element_type) {
        case B200FEM_TET4: fam = kFamTet4, nxe = 4; break;
        case B200FEM_HEX8: fam = kFamHex8, nxe = 8; break;
        case B200FEM_TET10: fam = kFamTet10, nxe = 10; break;
        case B200FEM_MACRO_TET4: fam = kFamMacroTet4, nxe = 10; break;
        default: set_error("packed laplacian: tet4 / hex8 / tet10 / macro_tet4 only"); return nullptr;
    }
    if (epp <= 0 || n_packs != (ne + epp - 1) / epp) {
        set_error("packed laplacian: n_packs does not match n_elements / n_elements_per_pack");
        return nullptr;
    }
    // validate the format: ranges inside [0, nn], local indices inside the pack, ghosts inside [0, nn)
    ptrdiff_t max_local = 0;
    for (ptrdiff_t p = 0; p < n_packs; ++p) {
        const ptrdiff_t n_owned = owned_nodes_ptr[p + 1] - owned_nodes_ptr[p], n_ghost = ghost_ptr[p + 1] - ghost_ptr[p];
        if (owned_nodes_ptr[p] < 0 || n_owned < 0 || owned_nodes_ptr[p + 1] > nn || n_ghost < 0 || n_shared[p] < 0 || n_shared[p] > n_owned) {
            set_error("packed laplacian: inconsistent owned / ghost ranges");
            return nullptr;
        }
        for (ptrdiff_t k = ghost_ptr[p]; k < ghost_ptr[p + 1]; ++k)
            if (ghost_idx[k] < 0 || ghost_idx[k] >= nn) {
                set_error("packed laplacian: ghost index out of range");
                return nullptr;
            }
        const ptrdiff_t n_local = n_owned + n_ghost;
        max_local = std::max(max_local, n_local);
        const ptrdiff_t e0 = p * epp, e1 = std::min(ne, e0 + epp);
        for (int v = 0; v < nxe; ++v)
            for (ptrdiff_t e = e0; e < e1; ++e)
                if ((ptrdiff_t)elements[v][e] >= n_local) {
                    set_error("packed laplacian: local node index outside its pack");
                    return nullptr;
                }
    }
    if (max_nodes < max_local) max_nodes = max_local;
    auto *op = new PackedLaplacian();
    op->fam = fam;
    op->elem_type = element_type;
    op->real_type = real_type;
    op->ne = ne;
    op->nn = nn;
    op->nxe = nxe;
    op->geom_bytes_per_elem = 24;
    op->n_packs = n_packs;
    op->epp = epp;
    op->max_nodes = max_nodes;
    std::vector<uint16_t> li((size_t)nxe * ne);
    for (int v = 0; v < nxe; ++v) std::copy(elements[v], elements[v] + ne, li.begin() + (size_t)v * ne);
    std::vector<float> f((size_t)6 * ne);
    for (ptrdiff_t e = 0; e < ne; ++e)
        for (int k = 0; k < 6; ++k) f[(size_t)k * ne + e] = fff[e * 6 + k];
    std::vector<int64_t> onp(owned_nodes_ptr, owned_nodes_ptr + n_packs + 1), gp(ghost_ptr, ghost_ptr + n_packs + 1);
    std::vector<int32_t> ns((size_t)n_packs);
    for (ptrdiff_t p = 0; p < n_packs; ++p) ns[(size_t)p] = (int32_t)n_shared[p];
    if (op->lidx.upload(li.data(), li.size()) || op->fff.upload(f.data(), f.size()) || op->owned_nodes_ptr.upload(onp.data(), onp.size()) ||
        op->ghost_ptr.upload(gp.data(), gp.size()) || op->n_shared.upload(ns.data(), ns.size()) ||
        op->ghost_idx.upload(ghost_idx, (size_t)ghost_ptr[n_packs])) {
        delete op;
        return nullptr;
    }
    return op;
}

int packed_laplacian_set_two_pass(OpBase *base, ptrdiff_t n_rows, const ptrdiff_t *ptr, const ptrdiff_t *idx, const int32_t *dest) {
    auto *op = dynamic_cast<PackedLaplacian *>(base);
    if (!op) B2_FAIL("set_two_pass: not a packed laplacian");
    const ptrdiff_t ng = (ptrdiff_t)op->ghost_idx.n;
    if (n_rows < 0 || (n_rows > 0 && (!ptr || !idx || !dest))) B2_FAIL("set_two_pass: NULL arrays");
    if ((size_t)2 * op->max_nodes * op->real_type > 200 * 1024)
        B2_FAIL("set_two_pass: a pack does not fit shared memory (the two-pass form stages the pack), use fewer elements per pack");
    // every ghost entry must be reduced exactly once, into the node it stands for
    if ((n_rows ? ptr[n_rows] : 0) != ng) B2_FAIL("set_two_pass: the reduce lists do not cover the ghost entries");
    std::vector<int32_t> host_ghost((size_t)ng);
    if (ng) B2_CUDA(cudaMemcpy(host_ghost.data(), op->ghost_idx.ptr, (size_t)ng * sizeof(int32_t), cudaMemcpyDeviceToHost));
    std::vector<uint8_t> seen((size_t)ng, 0);
    for (ptrdiff_t r = 0; r < n_rows; ++r) {
        if (dest[r] < 0 || dest[r] >= op->nn || ptr[r] > ptr[r + 1]) B2_FAIL("set_two_pass: bad row");
        for (ptrdiff_t j = ptr[r]; j < ptr[r + 1]; ++j) {
            if (idx[j] < 0 || idx[j] >= ng || seen[(size_t)idx[j]] || host_ghost[(size_t)idx[j]] != dest[r])
                B2_FAIL("set_two_pass: reduce lists inconsistent with ghost_idx");
            seen[(size_t)idx[j]] = 1;
        }
    }
    std::vector<int64_t> p64(ptr, ptr + (n_rows ? n_rows + 1 : 0)), i64(idx, idx + ng);
    if (p64.empty()) p64.push_back(0);
    if (op->reduce_ptr.upload(p64.data(), p64.size()) || op->reduce_idx.upload(i64.data(), i64.size()) ||
        op->reduce_dest.upload(dest, (size_t)n_rows) || op->ghost_buf.alloc((size_t)std::max<ptrdiff_t>(ng, 1) * op->real_type))
        return B200FEM_FAILURE;
    op->n_reduce_rows = n_rows;
    return B200FEM_SUCCESS;
}

}  // namespace b200fem

struct b200fem_packed_mesh_s {
    b200fem::PackedMeshData d;
};

extern "C" {

int b200fem_packed_mesh_create(b200fem_packed_mesh_t *pm, int nxe, ptrdiff_t ne, ptrdiff_t nn, const int32_t *const *elements,
                               ptrdiff_t n_elements_per_pack) {
    using namespace b200fem;
    if (!pm) B2_FAIL("packed mesh: NULL handle");
    *pm = nullptr;
    if (ne > 0 && !elements) B2_FAIL("packed mesh: elements is NULL");
    auto *h = new b200fem_packed_mesh_s();
    if (build_packed_mesh(h->d, nxe, ne, nn, elements, n_elements_per_pack)) {
        delete h;
        return B200FEM_FAILURE;
    }
    *pm = h;
    return B200FEM_SUCCESS;
}

int b200fem_packed_mesh_destroy(b200fem_packed_mesh_t pm) {
    delete pm;
    return B200FEM_SUCCESS;
}

int b200fem_packed_mesh_sizes(b200fem_packed_mesh_t pm, ptrdiff_t *n_packs, ptrdiff_t *n_elements_per_pack, ptrdiff_t *max_nodes_per_pack,
                              ptrdiff_t *n_ghost_entries) {
    using namespace b200fem;
    if (!pm) B2_FAIL("packed mesh: NULL handle");
    if (n_packs) *n_packs = pm->d.n_packs;
    if (n_elements_per_pack) *n_elements_per_pack = pm->d.n_elements_per_pack;
    if (max_nodes_per_pack) *max_nodes_per_pack = pm->d.max_nodes_per_pack;
    if (n_ghost_entries) *n_ghost_entries = (ptrdiff_t)pm->d.ghost_idx.size();
    return B200FEM_SUCCESS;
}

int b200fem_packed_mesh_arrays(b200fem_packed_mesh_t pm, const uint16_t *const **elements, const ptrdiff_t **owned_nodes_ptr,
                               const ptrdiff_t **n_shared, const ptrdiff_t **ghost_ptr, const int32_t **ghost_idx, const int32_t **new_to_old) {
    using namespace b200fem;
    if (!pm) B2_FAIL("packed mesh: NULL handle");
    if (elements) *elements = pm->d.row_ptrs.data();
    if (owned_nodes_ptr) *owned_nodes_ptr = pm->d.owned_nodes_ptr.data();
    if (n_shared) *n_shared = pm->d.n_shared.data();
    if (ghost_ptr) *ghost_ptr = pm->d.ghost_ptr.data();
    if (ghost_idx) *ghost_idx = pm->d.ghost_idx.data();
    if (new_to_old) *new_to_old = pm->d.new_to_old.data();
    return B200FEM_SUCCESS;
}

int b200fem_packed_mesh_ghost_reduce(b200fem_packed_mesh_t pm, ptrdiff_t *n_rows, const ptrdiff_t **ghost_reduce_ptr,
                                     const ptrdiff_t **ghost_reduce_idx, const int32_t **ghost_reduce_dest) {
    using namespace b200fem;
    if (!pm) B2_FAIL("packed mesh: NULL handle");
    if (n_rows) *n_rows = (ptrdiff_t)pm->d.ghost_reduce_dest.size();
    if (ghost_reduce_ptr) *ghost_reduce_ptr = pm->d.ghost_reduce_ptr.data();
    if (ghost_reduce_idx) *ghost_reduce_idx = pm->d.ghost_reduce_idx.data();
    if (ghost_reduce_dest) *ghost_reduce_dest = pm->d.ghost_reduce_dest.data();
    return B200FEM_SUCCESS;
}

}  // extern "C"
