// sshex8 (semi-structured hex8) Laplacian, matrix-free -- the north-star kernel.
//
// Replaces affine_sshex8_laplacian_stencil_apply_fff (operators/sshex8/sshex8_laplacian.cpp:698-769 of
// the reference) and affine_sshex8_laplacian_diag (:333-435); supersedes the CUDA twin
// cu_affine_sshex8_laplacian_apply_kernel_warp (operators/sshex8/cuda/cu_sshex8_laplacian_warp.hpp:4-99).
//
// Kernel shape (per CTA iteration = one batch of EPB macro elements, persistent grid):
//   1. gather   : all threads copy the (L+1)^3 lattice values of each element into a shared-memory tile,
//                 layout tile[z][slot][y][x]  (row stride N = L+1 doubles, odd => conflict-free LDS.64);
//   2. compute  : thread (slot, y) owns one x-row and marches over z (sshex8_laplacian_core.cuh): per
//                 input plane 3 rows are read from shared memory and 3 output planes are advanced
//                 (225 FMA per 27 shared loads); the finished output plane overwrites the input plane
//                 one step behind (in place, one barrier per step);
//   3. scatter  : interior nodes (owned by exactly one element) are written with plain stores /
//                 read-modify-write, skeleton nodes with RED.ADD.
//
// Connectivity in HBM comes in two layouts, chosen at plan creation:
//   MODE_EXPLICIT : int32 AoS [ne][(L+1)^3]   (one coalesced 4 B/node stream, any numbering)
//   MODE_BOXES    : 27 affine "entity boxes" per element (8 corners, 12 edges, 6 faces, 1 interior),
//                   each (base, sx, sy, sz) so that id(x,y,z) = base + sx x + sy y + sz z inside the
//                   box: 432 B per element instead of 2916 B at L = 8. Every numbering that orders the
//                   nodes of an edge / face / interior affinely in the local lattice (the reference's
//                   generator, mesh/sfem_hex8_mesh_graph.cpp:447-735, does) is detected; anything else
//                   (e.g. the hierarchical renumbering) falls back to MODE_EXPLICIT.
#include <string.h>

#include <algorithm>
#include <utility>
#include <vector>

#include "common.cuh"
#include "sshex8_boxes.cuh"
#include "sshex8_laplacian_core.cuh"
#include "sshex8_laplacian_stream.cuh"
#include "sshex8_laplacian_tma.cuh"

namespace b200fem {

static constexpr int kModeExplicit = 0;
static constexpr int kModeBoxes = 1;


template <int L>
__device__ __forceinline__ int box_of(int x, int y, int z) {
    const int cx = (x == 0) ? 0 : (x == L ? 2 : 1);
    const int cy = (y == 0) ? 0 : (y == L ? 2 : 1);
    const int cz = (z == 0) ? 0 : (z == L ? 2 : 1);
    return cx + 3 * cy + 9 * cz;
}

template <typename T, int L, int BT, int MODE, bool OVERWRITE, bool INTERIOR_EXCLUSIVE>
__global__ void __launch_bounds__(BT)
    sshex8_laplacian_kernel(const ptrdiff_t ne, const int32_t *__restrict__ idx_aos,
                            const BoxDesc *__restrict__ boxes, const float *__restrict__ fff,
                            const T *__restrict__ u, T *__restrict__ y) {
    constexpr int N = L + 1;
    constexpr int NXE = N * N * N;
    constexpr int EPB = BT / N;           // elements per batch
    constexpr int ZS = EPB * N * N;       // tile plane stride
    constexpr int NCOMPUTE = EPB * N;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    // [EPB][27] box descriptors first (16 B aligned for LDS.128), then the tile
    BoxDesc *sbox = reinterpret_cast<BoxDesc *>(smem_raw);
    T *tile = reinterpret_cast<T *>(smem_raw + (MODE == kModeBoxes ? EPB * 27 * sizeof(BoxDesc) : 0));

    const int tid = threadIdx.x;
    const ptrdiff_t nbatches = (ne + EPB - 1) / EPB;

    for (ptrdiff_t batch = blockIdx.x; batch < nbatches; batch += gridDim.x) {
        const ptrdiff_t e0 = batch * EPB;
        const int nslots = (int)((ne - e0) < (ptrdiff_t)EPB ? (ne - e0) : (ptrdiff_t)EPB);

        if (MODE == kModeBoxes) {
            const int32_t *src = reinterpret_cast<const int32_t *>(boxes + e0 * 27);
            int32_t *dst = reinterpret_cast<int32_t *>(sbox);
            for (int i = tid; i < nslots * 27 * 4; i += BT) dst[i] = ldg_stream(&src[i]);
            __syncthreads();
        }

        // ---- 1. gather
        for (int slot = 0; slot < nslots; ++slot) {
            const ptrdiff_t e = e0 + slot;
#pragma unroll 4
            for (int n = tid; n < NXE; n += BT) {
                const int z = n / (N * N), rem = n - z * (N * N);
                const int yy = rem / N, x = rem - yy * N;
                int32_t id;
                if (MODE == kModeExplicit) {
                    id = ldg_stream(&idx_aos[e * NXE + n]);
                } else {
                    const BoxDesc d = sbox[slot * 27 + box_of<L>(x, yy, z)];
                    id = d.base + d.sx * x + d.sy * yy + d.sz * z;
                }
                tile[z * ZS + (slot * N + yy) * N + x] = u[id];
            }
        }
        __syncthreads();

        // ---- 2. compute (in place)
        {
            const bool active = tid < NCOMPUTE && (tid / N) < nslots;
            Sshex8LaplacianRow<T, L> row;
            const int slot = tid / N, yrow = tid - slot * N;
            if (active) {
                T f[6];
#pragma unroll
                for (int k = 0; k < 6; ++k) f[k] = (T)__ldg(&fff[(e0 + slot) * 6 + k]);
                row.init(f, yrow);
            }
            T *base = tile + slot * N * N;
            for (int zp = 0; zp <= L; ++zp) {
                if (active) row.step(zp, base + zp * ZS, N);
                __syncthreads();
                if (active) row.write_and_rotate(zp, base + (zp > 0 ? zp - 1 : 0) * ZS, N);
            }
            if (active) row.write_last(base + L * ZS, N);
        }
        __syncthreads();

        // ---- 3. scatter
        for (int slot = 0; slot < nslots; ++slot) {
            const ptrdiff_t e = e0 + slot;
#pragma unroll 4
            for (int n = tid; n < NXE; n += BT) {
                const int z = n / (N * N), rem = n - z * (N * N);
                const int yy = rem / N, x = rem - yy * N;
                const T val = tile[z * ZS + (slot * N + yy) * N + x];
                if (MODE == kModeExplicit) {
                    const int32_t id = __ldg(&idx_aos[e * NXE + n]);
                    red_add(&y[id], val);
                } else {
                    const int b = box_of<L>(x, yy, z);
                    const BoxDesc d = sbox[slot * 27 + b];
                    const int32_t id = d.base + d.sx * x + d.sy * yy + d.sz * z;
                    if (INTERIOR_EXCLUSIVE && b == 13) {
                        if (OVERWRITE)
                            y[id] = val;
                        else
                            y[id] += val;
                    } else {
                        red_add(&y[id], val);
                    }
                }
            }
        }
        __syncthreads();
    }
}

// diag += diag(A): one thread per (element, lattice node)
template <typename T>
__global__ void __launch_bounds__(256)
    sshex8_laplacian_diag_kernel(const int L, const ptrdiff_t ne, const int32_t *__restrict__ idx_aos,
                                 const BoxDesc *__restrict__ boxes, const float *__restrict__ fff, T *__restrict__ diag) {
    const int N = L + 1, NXE = N * N * N;
    const ptrdiff_t total = ne * NXE;
    for (ptrdiff_t i = (ptrdiff_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (ptrdiff_t)gridDim.x * blockDim.x) {
        const ptrdiff_t e = i / NXE;
        const int n = (int)(i - e * NXE);
        const int z = n / (N * N), rem = n - z * (N * N);
        const int yy = rem / N, x = rem - yy * N;
        int32_t id;
        if (idx_aos) {
            id = idx_aos[i];
        } else {
            const int cx = (x == 0) ? 0 : (x == L ? 2 : 1), cy = (yy == 0) ? 0 : (yy == L ? 2 : 1),
                      cz = (z == 0) ? 0 : (z == L ? 2 : 1);
            const BoxDesc d = boxes[e * 27 + cx + 3 * cy + 9 * cz];
            id = d.base + d.sx * x + d.sy * yy + d.sz * z;
        }
        const T h = T(1) / T(L);
        const T Sxx = (T)fff[e * 6 + 0] * h, Sxy = (T)fff[e * 6 + 1] * h, Sxz = (T)fff[e * 6 + 2] * h,
                Syy = (T)fff[e * 6 + 3] * h, Syz = (T)fff[e * 6 + 4] * h, Szz = (T)fff[e * 6 + 5] * h;
        const T m0 = T(2.0 / 3.0);
        const T sx = (x == 0) ? T(-1) : (x == L ? T(1) : T(0));
        const T sy = (yy == 0) ? T(-1) : (yy == L ? T(1) : T(0));
        const T sz = (z == 0) ? T(-1) : (z == L ? T(1) : T(0));
        const T hx = (sx != T(0)) ? T(0.5) : T(1), hy = (sy != T(0)) ? T(0.5) : T(1), hz = (sz != T(0)) ? T(0.5) : T(1);
        T d = (Sxx + Syy + Szz) * T(2) * m0 * m0 * hx * hy * hz;
        d += T(0.5) * m0 * (Sxy * sx * sy * hz + Sxz * sx * sz * hy + Syz * sy * sz * hx);
        red_add(&diag[id], d);
    }
}

// ---------------------------------------------------------------------------------------------

struct Sshex8Laplacian final : OpBase {
    int mode = kModeExplicit;
    bool interior_exclusive = false;
    int32_t shared_prefix = 0;          // first skeleton range is [0, shared_prefix) (informational)
    std::vector<std::pair<int64_t, int64_t>> zero_ranges;  // skeleton id ranges = what an overwriting apply zeroes
    DeviceArray<int32_t> idx_aos;       // MODE_EXPLICIT: [ne][nxe]
    DeviceArray<BoxDesc> boxes;         // MODE_BOXES: [ne][27]
    DeviceArray<float> fff;             // [ne][6]
    DeviceArray<Sshex8Header> headers;  // MODE_BOXES: [ne] boxes + fff, what the streaming kernel reads
    DeviceArray<Sshex8TmaHeader> tma_headers;  // MODE_BOXES with contiguous boxes: what the TMA kernel reads
    bool tma_ok = false;                // every entity box is a contiguous id run (and descriptors fit int16)
    // 1 = entity-box TMA kernel where it applies. Off by default in round 1: measured at C3 it is on par with the
    // plane-streaming kernel (fp64 1.05 vs 0.98 ms, profiles/r01d_*): its warps spend half of every group waiting
    // for their own loads / stores because one buffer per element is all that fits next to 250 registers x 8 warps.
    int use_tma = 0;
    int tma_cfg = 0;                    // 0 = edges / corners per node, 1 = all 27 boxes as bulk copies
    int ctas_per_sm = 4;
    int stream_cfg = 0;                 // which team shape / ring depth (see launch_stream)
    int use_stream = 1;                 // 0 = the batch kernel below (kept for explicit connectivity)

    int reserve_sms = 0;
    // Warp-specialised kernel (sshex8_laplacian_ws.cuh) where it applies: level 8, compact numbering, 16-byte aligned
    // x. 1 = where it is the faster one (fp64: 0.90 against 0.97 - 1.04 ms at C3; at fp32 the z-unrolled plane-streaming
    // kernel wins, 0.66 against 0.79 ms), 2 = whenever it applies, 0 = never. The plane-streaming kernel stays the path
    // for everything else (and for the ne % 3 left-over elements).
    int use_ws = 1;
    int ws_cfg = 11;  // 11 = the two-barrier kernel with static accumulator phases, ring depth 3 (0 - 4: the first ws kernel)
    ptrdiff_t range_e0 = 0, range_e1 = -1;  // element range of the stream launch in flight (-1: the whole mesh)

    template <typename T, int L, int TW, int NTEAM, int D, bool PLAIN>
    int launch_stream_cfg(const T *x, T *y, cudaStream_t s) {
        const ptrdiff_t e_first = range_e1 < 0 ? 0 : range_e0;
        const ptrdiff_t ne = (range_e1 < 0 ? this->ne : range_e1) - e_first;
        const Sshex8Header *hdr = headers.ptr + e_first;
        using Cfg = Sshex8StreamCfg<T, L, TW, NTEAM, D>;
        auto kern = sshex8_laplacian_stream_kernel<T, L, TW, NTEAM, D, PLAIN, false>;
        auto tail = sshex8_laplacian_stream_kernel<T, L, TW, NTEAM, D, PLAIN, true>;
        static SmemOptIn opt_main, opt_tail;
        if (opt_main.ensure(kern, Cfg::smem_bytes) || opt_tail.ensure(tail, Cfg::smem_bytes)) return B200FEM_FAILURE;
        // full groups in the main launch, the ne % EPT left-over elements in a one-CTA clean-up launch
        const ptrdiff_t ngroups = ne / Cfg::EPT;
        const ptrdiff_t ne_main = ngroups * Cfg::EPT;
        if (ngroups > 0) {
            // `reserve_sms` SMs stay free for concurrently running kernels (the NCCL send / recv of an overlapped halo
            // exchange): this kernel's CTAs hold a whole SM each until the launch ends
            const int grid = (int)std::min<ptrdiff_t>((ngroups + NTEAM - 1) / NTEAM, (ptrdiff_t)std::max(1, sm_count() - reserve_sms));
            kern<<<grid, Cfg::threads, Cfg::smem_bytes, s>>>(ne_main, hdr, x, y);
        }
        if (ne_main < ne) tail<<<1, Cfg::threads, Cfg::smem_bytes, s>>>(ne - ne_main, hdr + ne_main, x, y);
        B2_CHECK_LAUNCH();
        return B200FEM_SUCCESS;
    }

    template <typename T>
    bool ws_applies(const T *x) const {
        return (use_ws == 2 || (use_ws == 1 && sizeof(T) == 8)) && level == 8 && interior_exclusive && range_e1 < 0 && (((uintptr_t)x) & 15) == 0 &&
               nn >= 64;
    }

    // warp-specialised kernel (sshex8_laplacian_ws.cu); a partial last group of 3 reads the padding headers (copies of
    // the last element, see make_sshex8_laplacian) and scatters only its valid slots
    template <typename T, bool PLAIN>
    int launch_ws(const T *x, T *y, cudaStream_t s) {
        const int max_ctas = std::max(1, sm_count() - reserve_sms);
        return sshex8_laplacian_ws_launch(sizeof(T) == 8 ? B200FEM_F64 : B200FEM_F32, PLAIN, ws_cfg, ne, nn, headers.ptr, x, y, max_ctas, s);
    }

    template <typename T, int L, bool PLAIN>
    int launch_tma(const T *x, T *y, cudaStream_t s) {
        constexpr int NWARP = 8;
        using Cfg = Sshex8TmaCfg<T, L, NWARP>;
        auto kern = tma_cfg == 1 ? sshex8_laplacian_tma_kernel<T, L, NWARP, PLAIN, false>
                                 : sshex8_laplacian_tma_kernel<T, L, NWARP, PLAIN, true>;
        static SmemOptIn opt[2];
        if (opt[tma_cfg == 1].ensure(kern, Cfg::smem_bytes)) return B200FEM_FAILURE;
        const ptrdiff_t ngroups = (ne + Cfg::EPW - 1) / Cfg::EPW;
        if (ngroups == 0) return B200FEM_SUCCESS;
        const int grid = (int)std::min<ptrdiff_t>((ngroups + NWARP - 1) / NWARP, (ptrdiff_t)sm_count());
        kern<<<grid, Cfg::threads, Cfg::smem_bytes, s>>>(ne, nn, tma_headers.ptr, x, y);
        B2_CHECK_LAUNCH();
        return B200FEM_SUCCESS;
    }

    // bulk copies need 16-byte aligned vectors; anything else takes the plane-streaming kernel
    template <typename T>
    bool tma_applies(const T *x, const T *y) const {
        return tma_ok && use_tma && level >= 4 && (((uintptr_t)x | (uintptr_t)y) & 15) == 0;
    }

    // team shapes per level: cfg 0 is the default, the others are tuning alternatives (bench sweeps them)
    template <typename T, int L, bool PLAIN>
    int launch_stream(const T *x, T *y, cudaStream_t s) {
        constexpr bool F64 = sizeof(T) == 8;
        if constexpr (L == 8) {
            if (ws_applies(x)) return launch_ws<T, PLAIN>(x, y, s);
            switch (stream_cfg) {
                default:
                case 0:  // fp32 runs the z-unrolled loop, which wants the ring slots static (depth 3 divides N = 9)
                    if constexpr (F64) return launch_stream_cfg<T, L, 1, 8, 6, PLAIN>(x, y, s);
                    else return launch_stream_cfg<T, L, 1, 8, 3, PLAIN>(x, y, s);
                case 1: return launch_stream_cfg<T, L, 2, 4, 7, PLAIN>(x, y, s);
                case 2: return launch_stream_cfg<T, L, 1, 8, 8, PLAIN>(x, y, s);
                case 3: return launch_stream_cfg<T, L, 2, 4, 5, PLAIN>(x, y, s);
                case 4: return launch_stream_cfg<T, L, 4, 2, F64 ? 7 : 9, PLAIN>(x, y, s);
                case 5: return launch_stream_cfg<T, L, 1, 6, 9, PLAIN>(x, y, s);
                case 6: return launch_stream_cfg<T, L, 1, 8, 9, PLAIN>(x, y, s);
                case 7: return launch_stream_cfg<T, L, 1, 8, 3, PLAIN>(x, y, s);
            }
        }
        else if constexpr (L == 4) {
            switch (stream_cfg) {
                default:
                case 0: return launch_stream_cfg<T, L, 1, 8, 5, PLAIN>(x, y, s);
                case 1: return launch_stream_cfg<T, L, 1, 16, 5, PLAIN>(x, y, s);
                case 2: return launch_stream_cfg<T, L, 2, 8, 5, PLAIN>(x, y, s);
            }
        }
        else {
            switch (stream_cfg) {
                default:
                case 0: return launch_stream_cfg<T, L, 1, 8, 3, PLAIN>(x, y, s);
                case 1: return launch_stream_cfg<T, L, 1, 16, 3, PLAIN>(x, y, s);
            }
        }
    }

    template <typename T, int L, int MODE, bool OVERWRITE, bool IE>
    int launch(const T *x, T *y, cudaStream_t s) {
        constexpr int BT = 64;
        constexpr int N = L + 1;
        constexpr int EPB = BT / N;
        const size_t smem = (size_t)EPB * N * N * N * sizeof(T) + (MODE == kModeBoxes ? (size_t)EPB * 27 * sizeof(BoxDesc) : 0);
        auto kern = sshex8_laplacian_kernel<T, L, BT, MODE, OVERWRITE, IE>;
        static SmemOptIn opt;
        if (opt.ensure(kern, smem)) return B200FEM_FAILURE;
        const ptrdiff_t nbatches = (ne + EPB - 1) / EPB;
        const int grid = (int)std::min<ptrdiff_t>(nbatches, (ptrdiff_t)sm_count() * ctas_per_sm);
        kern<<<grid, BT, smem, s>>>(ne, idx_aos.ptr, boxes.ptr, fff.ptr, x, y);
        B2_CHECK_LAUNCH();
        return B200FEM_SUCCESS;
    }

    template <typename T>
    int zero_skeleton(T *y, cudaStream_t s) {
        for (const auto &r : zero_ranges)
            B2_CUDA(cudaMemsetAsync(y + r.first, 0, (size_t)(r.second - r.first) * sizeof(T), s));
        return B200FEM_SUCCESS;
    }

    template <typename T, int L>
    int apply_L(const T *x, T *y, bool overwrite, cudaStream_t s) {
        if (mode == kModeBoxes && use_stream) {
            const bool tma = tma_applies(x, y);
            if (overwrite && interior_exclusive) {
                if (zero_skeleton(y, s)) return B200FEM_FAILURE;
                return tma ? launch_tma<T, L, true>(x, y, s) : launch_stream<T, L, true>(x, y, s);
            }
            if (overwrite) B2_CUDA(cudaMemsetAsync(y, 0, (size_t)nn * sizeof(T), s));
            return tma ? launch_tma<T, L, false>(x, y, s) : launch_stream<T, L, false>(x, y, s);
        }
        if (mode == kModeBoxes && interior_exclusive) {
            if (overwrite) {
                if (zero_skeleton(y, s)) return B200FEM_FAILURE;
                return launch<T, L, kModeBoxes, true, true>(x, y, s);
            }
            return launch<T, L, kModeBoxes, false, true>(x, y, s);
        }
        if (overwrite) B2_CUDA(cudaMemsetAsync(y, 0, (size_t)nn * sizeof(T), s));
        if (mode == kModeBoxes) return launch<T, L, kModeBoxes, false, false>(x, y, s);
        return launch<T, L, kModeExplicit, false, false>(x, y, s);
    }

    template <typename T, int L>
    int apply_range_L(const T *x, T *y, ptrdiff_t e0, ptrdiff_t e1, int mode, cudaStream_t s) {
        if (mode == 1) {
            if (interior_exclusive) {
                if (zero_skeleton(y, s)) return B200FEM_FAILURE;
            } else {
                B2_CUDA(cudaMemsetAsync(y, 0, (size_t)nn * sizeof(T), s));
            }
        }
        if (e1 == e0) return B200FEM_SUCCESS;
        if (L == 8 && ws_applies(x)) {
            // the warp-specialised kernel on the sub-range: a partial last group reads the headers that follow the range (or
            // the padding at the end of the array) for its missing slots and does not scatter them
            const int max_ctas = std::max(1, sm_count() - reserve_sms);
            return sshex8_laplacian_ws_launch(sizeof(T) == 8 ? B200FEM_F64 : B200FEM_F32, mode != 0 && interior_exclusive, ws_cfg, e1 - e0, nn,
                                              headers.ptr + e0, x, y, max_ctas, s);
        }
        range_e0 = e0;
        range_e1 = e1;
        const int rc = (mode != 0 && interior_exclusive) ? launch_stream<T, L, true>(x, y, s) : launch_stream<T, L, false>(x, y, s);
        range_e1 = -1;
        return rc;
    }

    // one launch of the warp-specialised kernel whose groups from e_dep / 3 on wait for the ghost block inside the kernel; one
    // SM stays free for the receiver half of the halo gather, which runs on another stream while this kernel holds the rest
    template <typename T>
    bool ghost_wait_applies(const T *x) const {
        return level == 8 && mode == kModeBoxes && use_stream && ws_cfg == 11 && ws_applies(x) && !tma_applies_flag();
    }
    bool tma_applies_flag() const { return use_tma != 0; }
    template <typename T>
    int apply_after_t(const T *x, T *y, ptrdiff_t e_dep, const unsigned long long *flag, unsigned long long want, cudaStream_t s) {
        if (!ghost_wait_applies(x)) B2_FAIL("apply_overwrite_after: needs the level-8 warp-specialised kernel (ws_cfg 11, compact numbering)");
        if (zero_skeleton(y, s)) return B200FEM_FAILURE;
        WsDependency dep;
        dep.flag = flag;
        dep.want = want;
        dep.first_group = (int)(e_dep / 3);
        const int max_ctas = std::max(1, sm_count() - std::max(1, reserve_sms));
        return sshex8_laplacian_ws_launch(sizeof(T) == 8 ? B200FEM_F64 : B200FEM_F32, true, ws_cfg, ne, nn, headers.ptr, x, y, max_ctas, s, &dep);
    }
    int apply_overwrite_after(const void *x, void *y, ptrdiff_t e_dep, const unsigned long long *flag, unsigned long long want,
                              cudaStream_t s) override {
        return real_type == B200FEM_F64 ? apply_after_t<double>((const double *)x, (double *)y, e_dep, flag, want, s)
                                        : apply_after_t<float>((const float *)x, (float *)y, e_dep, flag, want, s);
    }

    int apply_range(const void *x, void *y, ptrdiff_t e0, ptrdiff_t e1, int mode, cudaStream_t s) override {
        if (this->mode != kModeBoxes || !use_stream) B2_FAIL("apply_range: needs the boxes connectivity (stream kernel)");
        const bool f64 = real_type == B200FEM_F64;
        switch (level) {
            case 2: return f64 ? apply_range_L<double, 2>((const double *)x, (double *)y, e0, e1, mode, s) : apply_range_L<float, 2>((const float *)x, (float *)y, e0, e1, mode, s);
            case 4: return f64 ? apply_range_L<double, 4>((const double *)x, (double *)y, e0, e1, mode, s) : apply_range_L<float, 4>((const float *)x, (float *)y, e0, e1, mode, s);
            case 8: return f64 ? apply_range_L<double, 8>((const double *)x, (double *)y, e0, e1, mode, s) : apply_range_L<float, 8>((const float *)x, (float *)y, e0, e1, mode, s);
            default: B2_FAIL("sshex8 laplacian: level must be 2, 4 or 8");
        }
    }

    template <typename T>
    int apply_t(const T *x, T *y, bool overwrite, cudaStream_t s) {
        switch (level) {
            case 2: return apply_L<T, 2>(x, y, overwrite, s);
            case 4: return apply_L<T, 4>(x, y, overwrite, s);
            case 8: return apply_L<T, 8>(x, y, overwrite, s);
            default: B2_FAIL("sshex8 laplacian: level must be 2, 4 or 8 (other levels: not built yet)");
        }
    }

    int apply(const void *x, void *y, bool overwrite, cudaStream_t s) override {
        return real_type == B200FEM_F64 ? apply_t<double>((const double *)x, (double *)y, overwrite, s)
                                        : apply_t<float>((const float *)x, (float *)y, overwrite, s);
    }

    int diag(void *d, cudaStream_t s) override {
        const ptrdiff_t total = ne * nxe;
        const int grid = (int)std::min<ptrdiff_t>((total + 255) / 256, (ptrdiff_t)sm_count() * 16);
        if (real_type == B200FEM_F64)
            sshex8_laplacian_diag_kernel<double><<<grid, 256, 0, s>>>(level, ne, idx_aos.ptr, boxes.ptr, fff.ptr, (double *)d);
        else
            sshex8_laplacian_diag_kernel<float><<<grid, 256, 0, s>>>(level, ne, idx_aos.ptr, boxes.ptr, fff.ptr, (float *)d);
        B2_CHECK_LAUNCH();
        return B200FEM_SUCCESS;
    }

    double streamed_bytes(bool overwrite) const override {
        const double idx = (mode == kModeBoxes) ? ((tma_ok && use_tma && level >= 4) ? double(sizeof(Sshex8TmaHeader)) - 24.0 : 27.0 * sizeof(BoxDesc))
                                                : 4.0 * nxe;
        double vec = double(nn) * real_type * 2.0;                 // x read, y written
        if (!overwrite) vec += double(nn) * real_type;             // y read as well
        return double(ne) * (idx + 24.0) + vec;
    }
    int launches(bool overwrite) const override {
        // kernels + memset nodes. The warp-specialised kernel handles a partial last group itself; the plane-streaming kernel
        // (fp32, levels 2 / 4, unaligned vectors) adds a one-CTA clean-up launch when ne is not a multiple of its team size
        const bool ws = (use_ws == 2 || (use_ws == 1 && real_type == B200FEM_F64)) && level == 8 && interior_exclusive && nn >= 64;
        const int team = 32 / (level + 1);  // one-warp teams (the default stream configuration): 3 / 6 / 10 elements at level 8 / 4 / 2
        const int tail = (!ws && mode == kModeBoxes && use_stream && !use_tma && ne % team != 0) ? 1 : 0;
        return 1 + tail + (overwrite ? (int)std::max<size_t>(zero_ranges.size(), 1) : 0);
    }

    int set_option(const char *key, int value) override {
        if (!strcmp(key, "ctas_per_sm")) {
            if (value < 1 || value > 32) B2_FAIL("ctas_per_sm out of range");
            ctas_per_sm = value;
            return B200FEM_SUCCESS;
        }
        if (!strcmp(key, "stream_cfg")) { stream_cfg = value; return B200FEM_SUCCESS; }
        if (!strcmp(key, "use_stream")) { use_stream = value; return B200FEM_SUCCESS; }
        if (!strcmp(key, "use_tma")) { use_tma = value; return B200FEM_SUCCESS; }
        if (!strcmp(key, "use_ws")) { use_ws = value; return B200FEM_SUCCESS; }
        if (!strcmp(key, "ws_cfg")) { ws_cfg = value; return B200FEM_SUCCESS; }
        if (!strcmp(key, "reserve_sms")) {
            if (value < 0 || value >= sm_count()) B2_FAIL("reserve_sms out of range");
            reserve_sms = value;
            return B200FEM_SUCCESS;
        }
        if (!strcmp(key, "tma_cfg")) { tma_cfg = value; return B200FEM_SUCCESS; }
        if (!strcmp(key, "explicit_index")) {
            if (value && mode == kModeBoxes) {
                if (idx_aos.n == 0) B2_FAIL("explicit index was not kept");
                mode = kModeExplicit;
            } else if (!value && boxes.n != 0) {
                mode = kModeBoxes;
            }
            return B200FEM_SUCCESS;
        }
        return OpBase::set_option(key, value);
    }
    int get_option(const char *key, int *value) override {
        if (!strcmp(key, "mode")) { *value = mode; return B200FEM_SUCCESS; }
        if (!strcmp(key, "interior_exclusive")) { *value = interior_exclusive ? 1 : 0; return B200FEM_SUCCESS; }
        if (!strcmp(key, "shared_prefix")) { *value = shared_prefix; return B200FEM_SUCCESS; }
        if (!strcmp(key, "ctas_per_sm")) { *value = ctas_per_sm; return B200FEM_SUCCESS; }
        if (!strcmp(key, "stream_cfg")) { *value = stream_cfg; return B200FEM_SUCCESS; }
        if (!strcmp(key, "use_stream")) { *value = use_stream; return B200FEM_SUCCESS; }
        if (!strcmp(key, "use_tma")) { *value = use_tma; return B200FEM_SUCCESS; }
        if (!strcmp(key, "use_ws")) { *value = use_ws; return B200FEM_SUCCESS; }
        if (!strcmp(key, "ghost_wait")) {  // b200fem_op_apply_overwrite_after available (vector alignment is checked per call)
            *value = (level == 8 && mode == kModeBoxes && use_stream && ws_cfg == 11 && use_tma == 0 && interior_exclusive &&
                      (use_ws == 2 || (use_ws == 1 && real_type == B200FEM_F64)) && nn >= 64)
                             ? 1
                             : 0;
            return B200FEM_SUCCESS;
        }
        if (!strcmp(key, "tma_ok")) { *value = tma_ok ? 1 : 0; return B200FEM_SUCCESS; }
        return OpBase::get_option(key, value);
    }
};

// Per-element headers of the TMA kernel (sshex8_laplacian_tma.cuh). Returns false if some box is not a contiguous
// id run or a descriptor does not fit 16 bits: the plan then keeps the plane-streaming kernel.
template <typename T, int L>
static bool build_tma_headers_L(ptrdiff_t ne, const std::vector<BoxDesc> &boxes, const float *fff,
                                std::vector<Sshex8TmaHeader> &out) {
    constexpr int AL = 16 / (int)sizeof(T);
    const int lo[3] = {0, 1, L}, hi[3] = {0, L - 1, L};
    out.assign((size_t)ne, Sshex8TmaHeader{});
    for (ptrdiff_t e = 0; e < ne; ++e) {
        Sshex8TmaHeader &h = out[(size_t)e];
        for (int b = 0; b < 27; ++b) {
            const BoxDesc &d = boxes[(size_t)e * 27 + b];
            const int c[3] = {b % 3, (b / 3) % 3, b / 9};
            const int64_t st[3] = {d.sx, d.sy, d.sz};
            int64_t mn = d.base, mx = d.base;
            for (int a = 0; a < 3; ++a) {
                const int64_t v0 = st[a] * lo[c[a]], v1 = st[a] * hi[c[a]];
                mn += std::min(v0, v1);
                mx += std::max(v0, v1);
            }
            const int64_t count = tma_box_count<L>(b);
            if (mx - mn + 1 != count) return false;
            const int64_t shift = mn & (AL - 1);
            const int64_t A = (int64_t)tma_box_region<T, L>(b) + shift + ((int64_t)d.base - mn);
            if (A < INT16_MIN || A > INT16_MAX) return false;
            for (int a = 0; a < 3; ++a)
                if (st[a] < INT16_MIN || st[a] > INT16_MAX) return false;
            h.gmin[b] = (int32_t)mn;
            h.desc[b][0] = (int16_t)A;
            h.desc[b][1] = (int16_t)d.sx;
            h.desc[b][2] = (int16_t)d.sy;
            h.desc[b][3] = (int16_t)d.sz;
        }
        for (int k = 0; k < 6; ++k) h.fff[k] = fff[e * 6 + k];
    }
    return true;
}
template <typename T>
static bool build_tma_headers(int level, ptrdiff_t ne, const std::vector<BoxDesc> &boxes, const float *fff,
                              std::vector<Sshex8TmaHeader> &out) {
    switch (level) {
        case 2: return build_tma_headers_L<T, 2>(ne, boxes, fff, out);
        case 4: return build_tma_headers_L<T, 4>(ne, boxes, fff, out);
        case 8: return build_tma_headers_L<T, 8>(ne, boxes, fff, out);
        default: return false;
    }
}

OpBase *make_sshex8_laplacian(int level, ptrdiff_t ne, ptrdiff_t nn, const int32_t *const *elements, const float *fff,
                              int real_type) {
    if (level != 2 && level != 4 && level != 8) {  // the levels the apply kernels are instantiated for
        set_error("sshex8 laplacian: level must be 2, 4 or 8");
        return nullptr;
    }
    const int N = level + 1, nxe = N * N * N;
    for (int v = 0; v < nxe; ++v)
        for (ptrdiff_t e = 0; e < ne; ++e)
            if (elements[v][e] < 0 || elements[v][e] >= nn) {
                set_error("sshex8 laplacian: node index out of range");
                return nullptr;
            }
    auto *op = new Sshex8Laplacian();
    op->elem_type = B200FEM_SSHEX8;
    op->level = level;
    op->real_type = real_type;
    op->ne = ne;
    op->nn = nn;
    op->nxe = nxe;
    op->geom_bytes_per_elem = 24;

    std::vector<BoxDesc> boxes;
    bool excl = false;
    const bool affine = fit_boxes(level, ne, nn, elements, boxes, &excl, op->zero_ranges);
    int rc = 0;
    if (affine) {
        op->mode = kModeBoxes;
        op->interior_exclusive = excl;
        op->shared_prefix = (excl && !op->zero_ranges.empty() && op->zero_ranges[0].first == 0) ? (int32_t)op->zero_ranges[0].second : 0;
        rc |= op->boxes.upload(boxes.data(), boxes.size());
        std::vector<Sshex8Header> hdrs((size_t)ne);
        for (ptrdiff_t e = 0; e < ne; ++e) {
            Sshex8Header &h = hdrs[(size_t)e];
            for (int b = 0; b < 27; ++b) {
                const BoxDesc &d = boxes[(size_t)e * 27 + b];
                h.box[b][0] = d.base; h.box[b][1] = d.sx; h.box[b][2] = d.sy; h.box[b][3] = d.sz;
            }
            for (int k = 0; k < 6; ++k) h.fff[k] = fff[e * 6 + k];
            h.pad[0] = h.pad[1] = 0.f;
        }
        // two padding headers (copies of the last element): the ws kernel fetches headers by groups of 3
        if (ne > 0) {
            hdrs.push_back(hdrs.back());
            hdrs.push_back(hdrs.back());
        }
        rc |= op->headers.upload(hdrs.data(), hdrs.size());
        // TMA headers: every box must be a contiguous id run
        std::vector<Sshex8TmaHeader> th;
        const bool ok = (real_type == B200FEM_F64) ? build_tma_headers<double>(level, ne, boxes, fff, th)
                                                   : build_tma_headers<float>(level, ne, boxes, fff, th);
        if (ok) {
            rc |= op->tma_headers.upload(th.data(), th.size());
            op->tma_ok = true;
        }
    }
    // The explicit index is always kept: it is the fallback, serves the diag kernel when boxes are
    // absent, and lets tests run both connectivity modes on the same plan.
    {
        std::vector<int32_t> aos((size_t)ne * nxe);
        for (int v = 0; v < nxe; ++v) {
            const int32_t *row = elements[v];
            for (ptrdiff_t e = 0; e < ne; ++e) aos[(size_t)e * nxe + v] = row[e];
        }
        rc |= op->idx_aos.upload(aos.data(), aos.size());
    }
    rc |= op->fff.upload(fff, (size_t)ne * 6);
    if (rc) {
        delete op;
        return nullptr;
    }
    return op;
}

}  // namespace b200fem
