// sshex8 (semi-structured hex8) linear elasticity, matrix-free.
//
// Replaces affine_sshex8_linear_elasticity_apply_macro_adjugate
// (operators/sshex8/sshex8_linear_elasticity.cpp:275-424 of the reference: per macro element gather
// 3 x (L+1)^3 values, one 24x24 micro matrix, 512 dense 24x24 mat-vecs, atomic scatter) and
// affine_sshex8_linear_elasticity_diag / _block_diag_sym (:971-1256).
//
// Kernel shape (persistent grid; one CTA iteration = one batch of EPB macro elements):
//   1. gather  : the 3 x (L+1)^3 lattice values of every element of the batch go into a shared-memory
//                tile, SoA by component; the output tile is zeroed;
//   2. compute : the L^3 micro elements of a macro element are processed in 8 parity colours
//                ((x, y, z) mod 2); within a colour no two micro elements share a lattice node, so
//                (L/2)^3 threads per macro element update the output tile with plain shared-memory
//                read-modify-writes, one barrier per colour. The per-micro-element arithmetic is
//                hex8_le_core.cuh (gradient moments in the Haar basis: ~300 FMA + 144 add/sub, 15
//                per-macro-element constants in registers) -- the kernel is FP64-pipe bound, not HBM bound;
//   3. scatter : interior nodes (owned by exactly one element) with plain stores when y is being
//                overwritten, everything else with RED.ADD.
// Vectors are AoS [nn*3] (the reference's block_size-3 layout). Connectivity: 27 affine entity boxes per
// macro element when the numbering allows it (sshex8_boxes.cuh), explicit int32 AoS otherwise.
#include <string.h>

#include <algorithm>
#include <utility>
#include <vector>

#include "common.cuh"
#include "hex8_le_core.cuh"
#include "sshex8_boxes.cuh"

namespace b200fem {

static constexpr int kLEExplicit = 0;
static constexpr int kLEBoxes = 1;

template <int L, int EPB_ = 0>
struct Sshex8LECfg {
    static constexpr int N = L + 1;
    static constexpr int NXE = N * N * N;
    static constexpr int H = L / 2;            // micro elements per direction and colour
    static constexpr int TPE = H * H * H;      // threads per macro element
    // macro elements per CTA batch. L = 8: a 6-element tile fills the SM with ONE CTA whose 12 warps pass through gather,
    // compute and scatter together (memory idle while the FP64 pipe works and vice versa); 2- or 3-element tiles give 3 / 2
    // independent CTAs per SM that drift apart and overlap their phases (EPB_ = 0: the default of the level)
    static constexpr int EPB = EPB_ ? EPB_ : ((L == 8) ? 6 : (L == 4 ? 16 : 64));
    static constexpr int MIN_CTAS = (L == 8) ? 6 / EPB : 1;
    static constexpr int BT = TPE * EPB;
    // Shared-memory tile layout: index = z * PZ + y * N + xi(x) with the even x of a row first and the odd x behind
    // them (xi(x) = (x & 1) * XE + x / 2) and PZ = N^2 padded to 2 (mod 8). The threads of a half-warp are the (i, k)
    // pairs of one j (micro element (2i, 2j, 2k) + colour offset): consecutive i are consecutive doubles and k
    // advances by 2 PZ = 4 (mod 16) doubles, so the 16 8-byte accesses of a half-warp fall on 16 different bank
    // pairs. The dense layout with (i, j) half-warps put them on 2 (i + j): up to 6 wavefronts per request, and ncu
    // showed the kernel MIO-bound (5 bank conflicts per shared load; 5.1 -> 3.7 ms with 2-way conflicts, see DESIGN).
    static constexpr int XE = (N + 1) / 2;
    static constexpr int PZ = (N * N + 5) / 8 * 8 + 2;
    static constexpr int CS = N * PZ;  // component stride of a tile
    static_assert(PZ >= N * N && PZ % 8 == 2, "plane stride");
    __host__ __device__ static constexpr int xi(int x) { return (x & 1) * XE + (x >> 1); }
    static_assert(L % 2 == 0, "parity colouring needs an even level");
};

template <typename T, int L, int MODE, bool PLAIN_INTERIOR, int EPB_ = 0>
__global__ void __launch_bounds__(Sshex8LECfg<L, EPB_>::BT, Sshex8LECfg<L, EPB_>::MIN_CTAS)
    sshex8_le_apply_kernel(const ptrdiff_t ne, const int32_t *__restrict__ idx_aos, const BoxDesc *__restrict__ boxes,
                           const float *__restrict__ adjugate, const float *__restrict__ determinant, const T mu,
                           const T lambda, const T *__restrict__ u, T *__restrict__ y) {
    using Cfg = Sshex8LECfg<L, EPB_>;
    constexpr int N = Cfg::N, NXE = Cfg::NXE, H = Cfg::H, TPE = Cfg::TPE, EPB = Cfg::EPB, BT = Cfg::BT;
    constexpr int PZ = Cfg::PZ, CS = Cfg::CS;
    constexpr int ES = 3 * CS;  // values per element tile

    extern __shared__ __align__(16) unsigned char smem_raw[];
    BoxDesc *sbox = reinterpret_cast<BoxDesc *>(smem_raw);
    T *tile_u = reinterpret_cast<T *>(smem_raw + (MODE == kLEBoxes ? EPB * 27 * sizeof(BoxDesc) : 0));
    T *tile_o = tile_u + (size_t)EPB * ES;

    const int tid = threadIdx.x;
    const ptrdiff_t nbatches = (ne + EPB - 1) / EPB;

    for (ptrdiff_t batch = blockIdx.x; batch < nbatches; batch += gridDim.x) {
        const ptrdiff_t e0 = batch * EPB;
        const int nslots = (int)((ne - e0) < (ptrdiff_t)EPB ? (ne - e0) : (ptrdiff_t)EPB);

        if (MODE == kLEBoxes) {
            const int32_t *src = reinterpret_cast<const int32_t *>(boxes + e0 * 27);
            int32_t *dst = reinterpret_cast<int32_t *>(sbox);
            for (int i = tid; i < nslots * 27 * 4; i += BT) dst[i] = ldg_stream(&src[i]);
            __syncthreads();
        }

        // ---- 1. gather
        for (int i = tid; i < nslots * NXE; i += BT) {
            const int slot = i / NXE, n = i - slot * NXE;
            const int z = n / (N * N), rem = n - z * (N * N);
            const int yy = rem / N, x = rem - yy * N;
            int32_t id;
            if (MODE == kLEExplicit) {
                id = ldg_stream(&idx_aos[(e0 + slot) * NXE + n]);
            } else {
                const BoxDesc d = sbox[slot * 27 + sshex8_box_class(L, x, yy, z)];
                id = d.base + d.sx * x + d.sy * yy + d.sz * z;
            }
            const T *p = u + (ptrdiff_t)id * 3;
            const int tn = z * PZ + yy * N + Cfg::xi(x);
            T *tu = tile_u + slot * ES + tn, *to = tile_o + slot * ES + tn;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                tu[c * CS] = p[c];
                to[c * CS] = T(0);
            }
        }
        __syncthreads();

        // ---- 2. compute
        {
            const int slot = tid / TPE, t = tid - slot * TPE;
            const bool active = slot < nslots;
            const int j = t / (H * H), k = (t - j * H * H) / H, i = t - j * H * H - k * H;  // half-warp = (i, k) pairs
            Hex8LEConsts<T> K;
            if (active) {
                T A[9];
#pragma unroll
                for (int q = 0; q < 9; ++q) A[q] = (T)__ldg(&adjugate[(e0 + slot) * 9 + q]);
                K.init(A, (T)__ldg(&determinant[e0 + slot]), T(1) / T(L), mu, lambda);
            }
            const T *tu = tile_u + slot * ES;
            T *to = tile_o + slot * ES;
#pragma unroll 1
            for (int colour = 0; colour < 8; ++colour) {
                if (active) {
                    const int x0 = 2 * i + (colour & 1);
                    const int base = (2 * k + ((colour >> 2) & 1)) * PZ + (2 * j + ((colour >> 1) & 1)) * N;
                    const int xo[2] = {Cfg::xi(x0), Cfg::xi(x0 + 1)};
                    T ul[3][8], ol[3][8];
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
#pragma unroll
                        for (int v = 0; v < 8; ++v)
                            ul[c][v] = tu[c * CS + base + ((v >> 2) & 1) * PZ + ((v >> 1) & 1) * N + xo[v & 1]];
                    }
                    hex8_le_local_apply(K, ul, ol);
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
#pragma unroll
                        for (int v = 0; v < 8; ++v)
                            to[c * CS + base + ((v >> 2) & 1) * PZ + ((v >> 1) & 1) * N + xo[v & 1]] += ol[c][v];
                    }
                }
                __syncthreads();
            }
        }

        // ---- 3. scatter
        for (int i = tid; i < nslots * NXE; i += BT) {
            const int slot = i / NXE, n = i - slot * NXE;
            const int z = n / (N * N), rem = n - z * (N * N);
            const int yy = rem / N, x = rem - yy * N;
            int32_t id;
            bool plain = false;
            if (MODE == kLEExplicit) {
                id = __ldg(&idx_aos[(e0 + slot) * NXE + n]);
            } else {
                const int b = sshex8_box_class(L, x, yy, z);
                const BoxDesc d = sbox[slot * 27 + b];
                id = d.base + d.sx * x + d.sy * yy + d.sz * z;
                plain = PLAIN_INTERIOR && b == 13;
            }
            T *p = y + (ptrdiff_t)id * 3;
            const T *to = tile_o + slot * ES + z * PZ + yy * N + Cfg::xi(x);
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                if (plain)
                    __stcs(&p[c], to[c * CS]);
                else
                    red_add(&p[c], to[c * CS]);
            }
        }
        __syncthreads();
    }
}

// diag (BLOCK = false: 3 per node) / symmetric block diagonal (BLOCK = true: 6 per node); one thread per
// (element, lattice node)
template <typename T, bool BLOCK>
__global__ void __launch_bounds__(256)
    sshex8_le_diag_kernel(const int L, const ptrdiff_t ne, const int32_t *__restrict__ idx_aos,
                          const float *__restrict__ adjugate, const float *__restrict__ determinant, const T mu,
                          const T lambda, T *__restrict__ out) {
    const int N = L + 1, NXE = N * N * N;
    const ptrdiff_t total = ne * NXE;
    const T h = T(1) / T(L);
    for (ptrdiff_t i = (ptrdiff_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (ptrdiff_t)gridDim.x * blockDim.x) {
        const ptrdiff_t e = i / NXE;
        const int n = (int)(i - e * NXE);
        const int z = n / (N * N), rem = n - z * (N * N);
        const int yy = rem / N, x = rem - yy * N;
        const ptrdiff_t id = idx_aos[i];
        T A[9];
#pragma unroll
        for (int q = 0; q < 9; ++q) A[q] = (T)adjugate[e * 9 + q] * (h * h);
        const T det = (T)determinant[e] * (h * h * h);
        T blk[6];
        sshex8_le_node_block(A, det, mu, lambda, L, x, yy, z, blk);
        if (BLOCK) {
#pragma unroll
            for (int q = 0; q < 6; ++q) red_add(&out[id * 6 + q], blk[q]);
        } else {
            red_add(&out[id * 3 + 0], blk[0]);
            red_add(&out[id * 3 + 1], blk[3]);
            red_add(&out[id * 3 + 2], blk[5]);
        }
    }
}

struct Sshex8LinearElasticity final : OpBase {
    int mode = kLEExplicit;
    bool interior_exclusive = false;
    std::vector<std::pair<int64_t, int64_t>> zero_ranges;  // skeleton NODE ranges
    DeviceArray<int32_t> idx_aos;
    DeviceArray<BoxDesc> boxes;
    DeviceArray<float> adj, det;
    double mu = 1, lambda = 1;

    int epb = 0;  // L = 8 only: macro elements per CTA tile (1, 2, 3, 6); 0 = the measured best for the precision

    template <typename T, int L, int MODE, bool PLAIN>
    int launch(const T *x, T *y, cudaStream_t s) {
        if constexpr (L == 8) {
            // default: 1-element tiles in fp64 (2.53 ms at the C5 shard against 2.59 / 2.74 / 3.21 for 2 / 3 / 6), 3 in fp32
            // (2.25 ms against 2.51 / 2.41 / 2.44 for 1 / 2 / 6): profiles/r02n_le_sweep.jsonl
            const int e = epb ? epb : (sizeof(T) == 8 ? 1 : 3);
            if (e == 1) return launch_epb<T, L, MODE, PLAIN, 1>(x, y, s);
            if (e == 2) return launch_epb<T, L, MODE, PLAIN, 2>(x, y, s);
            if (e == 3) return launch_epb<T, L, MODE, PLAIN, 3>(x, y, s);
            return launch_epb<T, L, MODE, PLAIN, 6>(x, y, s);
        }
        return launch_epb<T, L, MODE, PLAIN, 0>(x, y, s);
    }

    template <typename T, int L, int MODE, bool PLAIN, int EPB_>
    int launch_epb(const T *x, T *y, cudaStream_t s) {
        using Cfg = Sshex8LECfg<L, EPB_>;
        const size_t smem = (size_t)Cfg::EPB * 2 * 3 * Cfg::CS * sizeof(T) + (MODE == kLEBoxes ? (size_t)Cfg::EPB * 27 * sizeof(BoxDesc) : 0);
        auto kern = sshex8_le_apply_kernel<T, L, MODE, PLAIN, EPB_>;
        static SmemOptIn opt;
        if (opt.ensure(kern, smem)) return B200FEM_FAILURE;
        const ptrdiff_t nbatches = (ne + Cfg::EPB - 1) / Cfg::EPB;
        const int per_sm = std::max(1, (int)((size_t)220 * 1024 / (smem + 1024)));
        const int grid = (int)std::min<ptrdiff_t>(nbatches, (ptrdiff_t)sm_count() * per_sm);
        kern<<<grid, Cfg::BT, smem, s>>>(ne, idx_aos.ptr, boxes.ptr, adj.ptr, det.ptr, (T)mu, (T)lambda, x, y);
        B2_CHECK_LAUNCH();
        return B200FEM_SUCCESS;
    }

    template <typename T, int L>
    int apply_L(const T *x, T *y, bool overwrite, cudaStream_t s) {
        if (mode == kLEBoxes) {
            if (overwrite && interior_exclusive) {
                for (const auto &r : zero_ranges)
                    B2_CUDA(cudaMemsetAsync(y + r.first * 3, 0, (size_t)(r.second - r.first) * 3 * sizeof(T), s));
                return launch<T, L, kLEBoxes, true>(x, y, s);
            }
            if (overwrite) B2_CUDA(cudaMemsetAsync(y, 0, (size_t)nn * 3 * sizeof(T), s));
            return launch<T, L, kLEBoxes, false>(x, y, s);
        }
        if (overwrite) B2_CUDA(cudaMemsetAsync(y, 0, (size_t)nn * 3 * sizeof(T), s));
        return launch<T, L, kLEExplicit, false>(x, y, s);
    }

    template <typename T>
    int apply_t(const T *x, T *y, bool overwrite, cudaStream_t s) {
        switch (level) {
            case 2: return apply_L<T, 2>(x, y, overwrite, s);
            case 4: return apply_L<T, 4>(x, y, overwrite, s);
            case 8: return apply_L<T, 8>(x, y, overwrite, s);
            default: B2_FAIL("sshex8 linear elasticity: level must be 2, 4 or 8");
        }
    }

    int apply(const void *x, void *y, bool overwrite, cudaStream_t s) override {
        return real_type == B200FEM_F64 ? apply_t<double>((const double *)x, (double *)y, overwrite, s)
                                        : apply_t<float>((const float *)x, (float *)y, overwrite, s);
    }

    template <bool BLOCK>
    int diag_impl(void *d, cudaStream_t s) {
        const ptrdiff_t total = ne * nxe;
        const int grid = (int)std::min<ptrdiff_t>((total + 255) / 256, (ptrdiff_t)sm_count() * 16);
        if (real_type == B200FEM_F64)
            sshex8_le_diag_kernel<double, BLOCK><<<grid, 256, 0, s>>>(level, ne, idx_aos.ptr, adj.ptr, det.ptr, mu, lambda, (double *)d);
        else
            sshex8_le_diag_kernel<float, BLOCK><<<grid, 256, 0, s>>>(level, ne, idx_aos.ptr, adj.ptr, det.ptr, (float)mu, (float)lambda, (float *)d);
        B2_CHECK_LAUNCH();
        return B200FEM_SUCCESS;
    }
    int diag(void *d, cudaStream_t s) override { return diag_impl<false>(d, s); }
    int block_diag_sym(void *d, cudaStream_t s) override { return diag_impl<true>(d, s); }

    double streamed_bytes(bool overwrite) const override {
        const double idx = (mode == kLEBoxes) ? 27.0 * sizeof(BoxDesc) : 4.0 * nxe;
        double vec = double(nn) * 3 * real_type * 2.0;
        if (!overwrite) vec += double(nn) * 3 * real_type;
        return double(ne) * (idx + 40.0) + vec;
    }
    int launches(bool overwrite) const override {
        return 1 + (overwrite ? (int)std::max<size_t>(zero_ranges.size(), 1) : 0);
    }
    int set_option(const char *key, int value) override {
        if (!strcmp(key, "epb")) {
            if (value != 0 && value != 1 && value != 2 && value != 3 && value != 6) B2_FAIL("epb must be 0 (default), 1, 2, 3 or 6");
            epb = value;
            return B200FEM_SUCCESS;
        }
        if (!strcmp(key, "explicit_index")) {
            if (value)
                mode = kLEExplicit;
            else if (boxes.n != 0)
                mode = kLEBoxes;
            return B200FEM_SUCCESS;
        }
        return OpBase::set_option(key, value);
    }
    int get_option(const char *key, int *value) override {
        if (!strcmp(key, "mode")) { *value = mode; return B200FEM_SUCCESS; }
        if (!strcmp(key, "interior_exclusive")) { *value = interior_exclusive ? 1 : 0; return B200FEM_SUCCESS; }
        return OpBase::get_option(key, value);
    }
};

OpBase *make_sshex8_linear_elasticity(int level, ptrdiff_t ne, ptrdiff_t nn, const int32_t *const *elements,
                                      const float *adjugate, const float *determinant, double mu, double lambda,
                                      int real_type) {
    if (level != 2 && level != 4 && level != 8) {
        set_error("sshex8 linear elasticity: level must be 2, 4 or 8");
        return nullptr;
    }
    const int N = level + 1, nxe = N * N * N;
    for (int v = 0; v < nxe; ++v)
        for (ptrdiff_t e = 0; e < ne; ++e)
            if (elements[v][e] < 0 || elements[v][e] >= nn) {
                set_error("sshex8 linear elasticity: node index out of range");
                return nullptr;
            }
    auto *op = new Sshex8LinearElasticity();
    op->elem_type = B200FEM_SSHEX8;
    op->level = level;
    op->real_type = real_type;
    op->block_size = 3;
    op->ne = ne;
    op->nn = nn;
    op->nxe = nxe;
    op->geom_bytes_per_elem = 40;
    op->mu = mu;
    op->lambda = lambda;

    std::vector<BoxDesc> boxes;
    bool excl = false;
    int rc = 0;
    if (fit_boxes(level, ne, nn, elements, boxes, &excl, op->zero_ranges)) {
        op->mode = kLEBoxes;
        op->interior_exclusive = excl;
        rc |= op->boxes.upload(boxes.data(), boxes.size());
    }
    {
        std::vector<int32_t> aos((size_t)ne * nxe);
        for (int v = 0; v < nxe; ++v) {
            const int32_t *row = elements[v];
            for (ptrdiff_t e = 0; e < ne; ++e) aos[(size_t)e * nxe + v] = row[e];
        }
        rc |= op->idx_aos.upload(aos.data(), aos.size());
    }
    rc |= op->adj.upload(adjugate, (size_t)ne * 9);
    rc |= op->det.upload(determinant, (size_t)ne);
    if (rc) {
        delete op;
        return nullptr;
    }
    return op;
}

}  // namespace b200fem
