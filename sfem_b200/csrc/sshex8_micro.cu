// Semi-structured operators evaluated micro element by micro element inside a shared-memory tile of the macro element:
//
//   * stored-element-matrix ("EM") operators: ss:em:Laplacian / ss:em:LinearElasticity of the reference
//     (frontend/ops/sfem_SemiStructuredEMLaplacian.cpp, sfem_SemiStructuredEMLinearElasticity.cpp; CPU kernels
//     operators/stencil/sshex8_stencil_element_matrix_apply.cpp:10-160, GPU twins GPUEMOp / GPUEMWarpOp,
//     operators/sshex8/cuda/cu_sshex8_elemental_matrix.cu:226-1053): ONE dense (8 bs) x (8 bs) matrix per macro element,
//     stored in HBM, applied to every one of its L^3 micro elements -- out_j += sum_i u_i M[i][j], VTK corner order,
//     components blocked [c0 x 8 | c1 x 8 | c2 x 8] (operators/sshex8/sshex8.hpp:44-86, operators/packed_elements.hpp:8-80);
//     with the builders of those matrices (sshex8_laplacian_element_matrix, operators/sshex8/sshex8_laplacian.cpp:595-634;
//     sshex8_linear_elasticity_element_matrix, sshex8_linear_elasticity.cpp:1258-1310). The element-matrix formulation of
//     the affine Laplacian (affine_sshex8_laplacian_apply, sshex8_laplacian.cpp:184-331) is this operator with the matrix
//     built from the macro FFF;
//   * the non-affine Laplacian sshex8_laplacian_apply (sshex8_laplacian.cpp:39-182): trilinear macro elements, FFF
//     evaluated at the 6 face-centre quadrature points (operators/hex8/hex8_quadrature.hpp:4-16) of every micro element.
//
// Kernel shape = the tile kernel of sshex8_linear_elasticity.cu: a persistent CTA takes batches of EPB macro elements;
// gather of the (L+1)^3 x bs lattice values into a conflict-free shared-memory tile, 8 parity colours of (L/2)^3 micro
// elements each (no two micro elements of a colour share a node: plain shared-memory read-modify-writes, one barrier per
// colour), scatter with RED.ADD. The per-macro-element data of the local operator (the matrix / the 8 corner
// coordinates) is staged in shared memory next to the tile and read as warp-wide broadcasts.
#include <string.h>

#include <algorithm>
#include <type_traits>
#include <vector>

#include "common.cuh"
#include "hex8_le_core.cuh"
#include "hex8_local.cuh"

namespace b200fem {

template <int L, int BS>
struct MicroCfg {
    static constexpr int N = L + 1;
    static constexpr int NXE = N * N * N;
    static constexpr int H = L / 2;
    static constexpr int TPE = H * H * H;  // threads per macro element = micro elements per colour
    // macro elements per CTA batch: >= 2 warps of work, tile of BS components within ~100 KB
    // small tiles, several independent CTAs per SM whose gather / compute / scatter phases overlap (the same finding as
    // for sshex8_le_apply_kernel): bs = 1: 2 elements, 128 threads, 3 CTAs (168 registers); bs = 3: 1 element, 64 threads, 4 CTAs
    static constexpr int EPB = (L == 8) ? (BS == 1 ? 2 : 1) : (L == 4 ? 16 : (BS == 1 ? 64 : 32));
    static constexpr int MIN_CTAS = (L == 8) ? (BS == 1 ? 3 : 4) : 1;
    static constexpr int BT = TPE * EPB;
    static constexpr int XE = (N + 1) / 2;
    static constexpr int PZ = (N * N + 5) / 8 * 8 + 2;  // plane stride = 2 (mod 8): see Sshex8LECfg
    static constexpr int CS = N * PZ;
    __host__ __device__ static constexpr int xi(int x) { return (x & 1) * XE + (x >> 1); }
};

// ---------------------------------------------------------------------------------------------------------------
// local operators
// ---------------------------------------------------------------------------------------------------------------
// stored matrix, VTK order; ul / ol come in lexicographic corner order (x + 2y + 4z).
// bs = 1: the 64 entries of the element's matrix are loaded ONCE into registers (a thread applies it to its 8 micro
// elements, one per colour): 8 + 16 shared-memory accesses per micro element instead of 64 + 24.
// bs = 3: 576 entries stay in shared memory and are read as 16-byte warp-wide broadcasts.
template <typename T, int BS>
struct MicroMatrixLocal {
    static constexpr int NB = 8 * BS;
    static constexpr int kSlotValues = NB * NB;
    struct Regs {
        T m[BS == 1 ? 64 : 1];
    };
    const T *matrices;  // [ne][NB * NB]
    __device__ __forceinline__ void stage(const ptrdiff_t e, T *slot, const int t, const int nt) const {
        const T *src = matrices + (size_t)e * kSlotValues;
        for (int i = t; i < kSlotValues; i += nt) slot[i] = ldg_stream(&src[i]);
    }
    __device__ __forceinline__ void load_regs(const T *slot, Regs &r) const {
        if constexpr (BS == 1) {
            // lexicographic on both sides: r.m[li * 8 + lj] = M[vtk(li)][vtk(lj)]
#pragma unroll
            for (int vi = 0; vi < 8; ++vi)
#pragma unroll
                for (int vj = 0; vj < 8; ++vj) r.m[kVtkToLex[vi] * 8 + kVtkToLex[vj]] = slot[vi * 8 + vj];
        }
    }
    __device__ __forceinline__ void apply(const Regs &r, const T *slot, int, int, int, const T (&ul)[BS][8], T (&ol)[BS][8]) const {
        if constexpr (BS == 1) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                T acc = ul[0][0] * r.m[j];
#pragma unroll
                for (int i = 1; i < 8; ++i) acc = le_fma(ul[0][i], r.m[i * 8 + j], acc);
                ol[0][j] = acc;
            }
        } else {
            using V = typename std::conditional<sizeof(T) == 8, double2, float4>::type;
            constexpr int VW = 16 / (int)sizeof(T);
#pragma unroll
            for (int c = 0; c < BS; ++c)
#pragma unroll
                for (int v = 0; v < 8; ++v) ol[c][v] = T(0);
            // out_j += u_i M[i][j]; i, j = component * 8 + VTK corner; a row is NB contiguous values
#pragma unroll
            for (int ci = 0; ci < BS; ++ci) {
#pragma unroll
                for (int vi = 0; vi < 8; ++vi) {
                    const T ui = ul[ci][kVtkToLex[vi]];
                    const V *row = reinterpret_cast<const V *>(slot + (ci * 8 + vi) * NB);
#pragma unroll
                    for (int q = 0; q < NB / VW; ++q) {
                        const V w = row[q];
                        const T *wv = reinterpret_cast<const T *>(&w);
#pragma unroll
                        for (int k = 0; k < VW; ++k) {
                            const int j = q * VW + k, cj = j / 8, vj = j % 8;
                            ol[cj][kVtkToLex[vj]] = le_fma(ui, wv[k], ol[cj][kVtkToLex[vj]]);
                        }
                    }
                }
            }
        }
    }
};

// sshex8_laplacian_apply: trilinear macro element, q6 quadrature per micro element
template <typename T>
struct NonAffineLaplacianLocal {
    static constexpr int kSlotValues = 24;  // corner coordinates x[8] | y[8] | z[8], VTK order
    struct Regs {};
    __device__ __forceinline__ void load_regs(const T *, Regs &) const {}
    const int32_t *idx_aos;                 // [ne][NXE]
    const float *points;                    // [3][nn]
    ptrdiff_t nn;
    int level, nxe;
    __device__ __forceinline__ void stage(const ptrdiff_t e, T *slot, const int t, const int nt) const {
        const int N = level + 1;
        for (int i = t; i < 24; i += nt) {
            const int d = i / 8, v = i - d * 8;
            const int lex = kVtkToLex[v];
            const int row = ((lex >> 2) & 1) * level * N * N + ((lex >> 1) & 1) * level * N + (lex & 1) * level;
            slot[i] = (T)points[(size_t)d * nn + idx_aos[(size_t)e * nxe + row]];
        }
    }
    // FFF = adj adj^T / det of the trilinear map at (qx, qy, qz): what hex8_fff computes (operators/hex8/hex8_inline_cpu.hpp:13-73)
    __device__ __forceinline__ static void fff_at(const T *X, const T qx, const T qy, const T qz, T fff[6]) {
        T J[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) J[k] = T(0);
#pragma unroll
        for (int v = 0; v < 8; ++v) {
            const int lex = kVtkToLex[v];
            const int bx = lex & 1, by = (lex >> 1) & 1, bz = (lex >> 2) & 1;
            const T sx = bx ? qx : T(1) - qx, sy = by ? qy : T(1) - qy, sz = bz ? qz : T(1) - qz;
            const T gx = (bx ? T(1) : T(-1)) * sy * sz, gy = sx * (by ? T(1) : T(-1)) * sz, gz = sx * sy * (bz ? T(1) : T(-1));
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                const T p = X[r * 8 + v];
                J[r * 3 + 0] = le_fma(p, gx, J[r * 3 + 0]);
                J[r * 3 + 1] = le_fma(p, gy, J[r * 3 + 1]);
                J[r * 3 + 2] = le_fma(p, gz, J[r * 3 + 2]);
            }
        }
        T A[9];
        A[0] = J[4] * J[8] - J[5] * J[7];
        A[1] = J[2] * J[7] - J[1] * J[8];
        A[2] = J[1] * J[5] - J[2] * J[4];
        A[3] = J[5] * J[6] - J[3] * J[8];
        A[4] = J[0] * J[8] - J[2] * J[6];
        A[5] = J[2] * J[3] - J[0] * J[5];
        A[6] = J[3] * J[7] - J[4] * J[6];
        A[7] = J[1] * J[6] - J[0] * J[7];
        A[8] = J[0] * J[4] - J[1] * J[3];
        const T idet = T(1) / (J[0] * A[0] + J[1] * A[3] + J[2] * A[6]);
        int k = 0;
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = i; j < 3; ++j) fff[k++] = (A[i * 3] * A[j * 3] + A[i * 3 + 1] * A[j * 3 + 1] + A[i * 3 + 2] * A[j * 3 + 2]) * idet;
    }
    __device__ __forceinline__ void apply(const Regs &, const T *slot, const int xi, const int yi, const int zi, const T (&ul)[1][8],
                                          T (&ol)[1][8]) const {
        const T h = T(1) / T(level);
        const T q6[6][3] = {{0, .5, .5}, {.5, 0, .5}, {.5, .5, 0}, {.5, .5, 1}, {.5, 1, .5}, {1, .5, .5}};
#pragma unroll
        for (int v = 0; v < 8; ++v) ol[0][v] = T(0);
#pragma unroll 1
        for (int q = 0; q < 6; ++q) {
            const T qx = q6[q][0], qy = q6[q][1], qz = q6[q][2];
            T m[6];
            fff_at(slot, le_fma(h, qx, T(xi) * h), le_fma(h, qy, T(yi) * h), le_fma(h, qz, T(zi) * h), m);
            // reference gradient of u at the quadrature point of the micro element (lexicographic corners)
            T g[3] = {0, 0, 0};
            T gx[8], gy[8], gz[8];
#pragma unroll
            for (int lex = 0; lex < 8; ++lex) {
                const int bx = lex & 1, by = (lex >> 1) & 1, bz = (lex >> 2) & 1;
                const T sx = bx ? qx : T(1) - qx, sy = by ? qy : T(1) - qy, sz = bz ? qz : T(1) - qz;
                gx[lex] = (bx ? T(1) : T(-1)) * sy * sz;
                gy[lex] = sx * (by ? T(1) : T(-1)) * sz;
                gz[lex] = sx * sy * (bz ? T(1) : T(-1));
                g[0] = le_fma(ul[0][lex], gx[lex], g[0]);
                g[1] = le_fma(ul[0][lex], gy[lex], g[1]);
                g[2] = le_fma(ul[0][lex], gz[lex], g[2]);
            }
            // hex8_sub_fff_0: fff = m_fff * h; weight 1/6
            const T w = h * T(1.0 / 6.0);
            const T t0 = w * (m[0] * g[0] + m[1] * g[1] + m[2] * g[2]);
            const T t1 = w * (m[1] * g[0] + m[3] * g[1] + m[4] * g[2]);
            const T t2 = w * (m[2] * g[0] + m[4] * g[1] + m[5] * g[2]);
#pragma unroll
            for (int lex = 0; lex < 8; ++lex) ol[0][lex] += gx[lex] * t0 + gy[lex] * t1 + gz[lex] * t2;
        }
    }
};

// ---------------------------------------------------------------------------------------------------------------
// the tile kernel
// ---------------------------------------------------------------------------------------------------------------
template <typename T, int L, int BS, class Local>
__global__ void __launch_bounds__(MicroCfg<L, BS>::BT, MicroCfg<L, BS>::MIN_CTAS)
    sshex8_micro_apply_kernel(const ptrdiff_t ne, const int32_t *__restrict__ idx_aos, const Local local,
                              const T *__restrict__ u, T *__restrict__ y) {
    using Cfg = MicroCfg<L, BS>;
    constexpr int N = Cfg::N, NXE = Cfg::NXE, H = Cfg::H, TPE = Cfg::TPE, EPB = Cfg::EPB, BT = Cfg::BT;
    constexpr int PZ = Cfg::PZ, CS = Cfg::CS, ES = BS * CS;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *tile_u = reinterpret_cast<T *>(smem_raw);
    T *tile_o = tile_u + (size_t)EPB * ES;
    T *slots = tile_o + (size_t)EPB * ES;  // [EPB][Local::kSlotValues]
    const int tid = threadIdx.x;
    const ptrdiff_t nbatches = (ne + EPB - 1) / EPB;
    for (ptrdiff_t batch = blockIdx.x; batch < nbatches; batch += gridDim.x) {
        const ptrdiff_t e0 = batch * EPB;
        const int nslots = (int)((ne - e0) < (ptrdiff_t)EPB ? (ne - e0) : (ptrdiff_t)EPB);
        // ---- per-element operator data + gather
        {
            const int slot = tid / TPE, t = tid - slot * TPE;
            if (slot < nslots) local.stage(e0 + slot, slots + (size_t)slot * Local::kSlotValues, t, TPE);
        }
        for (int i = tid; i < nslots * NXE; i += BT) {
            const int slot = i / NXE, n = i - slot * NXE;
            const int z = n / (N * N), rem = n - z * (N * N);
            const int yy = rem / N, x = rem - yy * N;
            const int32_t id = ldg_stream(&idx_aos[(e0 + slot) * NXE + n]);
            const T *p = u + (ptrdiff_t)id * BS;
            const int tn = z * PZ + yy * N + Cfg::xi(x);
#pragma unroll
            for (int c = 0; c < BS; ++c) {
                tile_u[slot * ES + c * CS + tn] = p[c];
                tile_o[slot * ES + c * CS + tn] = T(0);
            }
        }
        __syncthreads();
        // ---- 8 parity colours
        {
            const int slot = tid / TPE, t = tid - slot * TPE;
            const bool active = slot < nslots;
            const int j = t / (H * H), k = (t - j * H * H) / H, i = t - j * H * H - k * H;
            const T *tu = tile_u + slot * ES;
            T *to = tile_o + slot * ES;
            const T *sl = slots + (size_t)slot * Local::kSlotValues;
            typename Local::Regs regs;
            if (active) local.load_regs(sl, regs);
#pragma unroll 1
            for (int colour = 0; colour < 8; ++colour) {
                if (active) {
                    const int x0 = 2 * i + (colour & 1), y0 = 2 * j + ((colour >> 1) & 1), z0 = 2 * k + ((colour >> 2) & 1);
                    const int base = z0 * PZ + y0 * N;
                    const int xo[2] = {Cfg::xi(x0), Cfg::xi(x0 + 1)};
                    T ul[BS][8], ol[BS][8];
#pragma unroll
                    for (int c = 0; c < BS; ++c)
#pragma unroll
                        for (int v = 0; v < 8; ++v) ul[c][v] = tu[c * CS + base + ((v >> 2) & 1) * PZ + ((v >> 1) & 1) * N + xo[v & 1]];
                    local.apply(regs, sl, x0, y0, z0, ul, ol);
#pragma unroll
                    for (int c = 0; c < BS; ++c)
#pragma unroll
                        for (int v = 0; v < 8; ++v) to[c * CS + base + ((v >> 2) & 1) * PZ + ((v >> 1) & 1) * N + xo[v & 1]] += ol[c][v];
                }
                __syncthreads();
            }
        }
        // ---- scatter
        for (int i = tid; i < nslots * NXE; i += BT) {
            const int slot = i / NXE, n = i - slot * NXE;
            const int z = n / (N * N), rem = n - z * (N * N);
            const int yy = rem / N, x = rem - yy * N;
            const int32_t id = __ldg(&idx_aos[(e0 + slot) * NXE + n]);
            T *p = y + (ptrdiff_t)id * BS;
            const int tn = z * PZ + yy * N + Cfg::xi(x);
#pragma unroll
            for (int c = 0; c < BS; ++c) red_add(&p[c], tile_o[slot * ES + c * CS + tn]);
        }
        __syncthreads();
    }
}

// diag += diagonal of the assembled operator from the stored matrices: lattice node (x, y, z) is corner c of the micro
// elements (x - cx, y - cy, z - cz) that exist
template <typename T, int BS>
__global__ void __launch_bounds__(256)
    sshex8_em_diag_kernel(const int L, const ptrdiff_t ne, const int32_t *__restrict__ idx_aos, const T *__restrict__ matrices,
                          T *__restrict__ diag) {
    constexpr int NB = 8 * BS;
    const int N = L + 1, NXE = N * N * N;
    const ptrdiff_t total = ne * NXE;
    for (ptrdiff_t i = (ptrdiff_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (ptrdiff_t)gridDim.x * blockDim.x) {
        const ptrdiff_t e = i / NXE;
        const int n = (int)(i - e * NXE);
        const int z = n / (N * N), rem = n - z * (N * N);
        const int yy = rem / N, x = rem - yy * N;
        const T *M = matrices + (size_t)e * NB * NB;
        T acc[BS];
#pragma unroll
        for (int c = 0; c < BS; ++c) acc[c] = T(0);
#pragma unroll
        for (int v = 0; v < 8; ++v) {
            const int lex = kVtkToLex[v];
            const int mx = x - (lex & 1), my = yy - ((lex >> 1) & 1), mz = z - ((lex >> 2) & 1);
            if (mx < 0 || my < 0 || mz < 0 || mx >= L || my >= L || mz >= L) continue;
#pragma unroll
            for (int c = 0; c < BS; ++c) acc[c] += M[(c * 8 + v) * NB + (c * 8 + v)];
        }
        const ptrdiff_t id = idx_aos[i];
#pragma unroll
        for (int c = 0; c < BS; ++c) red_add(&diag[id * BS + c], acc[c]);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// element-matrix builders: the local operators applied to the unit vectors
// ---------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void sshex8_laplacian_em_kernel(const int level, const ptrdiff_t ne, const float *__restrict__ fff, T *__restrict__ out) {
    for (ptrdiff_t e = (ptrdiff_t)blockIdx.x * blockDim.x + threadIdx.x; e < ne; e += (ptrdiff_t)gridDim.x * blockDim.x) {
        const T h = T(1) / T(level);
        T f[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) f[k] = (T)fff[e * 6 + k] * h;  // hex8_sub_fff_0
        for (int vi = 0; vi < 8; ++vi) {
            T ul[8], ol[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) ul[q] = T(0);
            ul[kVtkToLex[vi]] = T(1);
            hex8_laplacian_local(f, ul, ol);
#pragma unroll
            for (int vj = 0; vj < 8; ++vj) out[(size_t)e * 64 + vi * 8 + vj] = ol[kVtkToLex[vj]];
        }
    }
}

template <typename T>
__global__ void sshex8_le_em_kernel(const int level, const ptrdiff_t ne, const float *__restrict__ adjugate,
                                    const float *__restrict__ determinant, const T mu, const T lambda, T *__restrict__ out) {
    for (ptrdiff_t e = (ptrdiff_t)blockIdx.x * blockDim.x + threadIdx.x; e < ne; e += (ptrdiff_t)gridDim.x * blockDim.x) {
        T A[9];
#pragma unroll
        for (int q = 0; q < 9; ++q) A[q] = (T)adjugate[e * 9 + q];
        Hex8LEConsts<T> K;
        K.init(A, (T)determinant[e], T(1) / T(level), mu, lambda);
        for (int i = 0; i < 24; ++i) {
            T ul[3][8], ol[3][8];
#pragma unroll
            for (int c = 0; c < 3; ++c)
#pragma unroll
                for (int q = 0; q < 8; ++q) ul[c][q] = T(0);
            ul[i / 8][kVtkToLex[i % 8]] = T(1);
            hex8_le_local_apply(K, ul, ol);
            for (int j = 0; j < 24; ++j) out[(size_t)e * 576 + i * 24 + j] = ol[j / 8][kVtkToLex[j % 8]];
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// plans
// ---------------------------------------------------------------------------------------------------------------
static int upload_aos(int level, ptrdiff_t ne, ptrdiff_t nn, const int32_t *const *elements, DeviceArray<int32_t> &idx) {
    const int N = level + 1, nxe = N * N * N;
    std::vector<int32_t> aos((size_t)ne * nxe);
    for (int v = 0; v < nxe; ++v) {
        const int32_t *row = elements[v];
        for (ptrdiff_t e = 0; e < ne; ++e) {
            if (row[e] < 0 || row[e] >= nn) B2_FAIL("sshex8 operator: node index out of range");
            aos[(size_t)e * nxe + v] = row[e];
        }
    }
    return idx.upload(aos.data(), aos.size());
}

template <typename T, int L, int BS, class Local>
static int launch_micro(ptrdiff_t ne, const int32_t *idx, const Local &local, const T *x, T *y, cudaStream_t s) {
    using Cfg = MicroCfg<L, BS>;
    const size_t smem = ((size_t)Cfg::EPB * 2 * BS * Cfg::CS + (size_t)Cfg::EPB * Local::kSlotValues) * sizeof(T);
    auto kern = sshex8_micro_apply_kernel<T, L, BS, Local>;
    static SmemOptIn opt;
    if (opt.ensure(kern, smem)) return B200FEM_FAILURE;
    if (ne == 0) return B200FEM_SUCCESS;
    const ptrdiff_t nbatches = (ne + Cfg::EPB - 1) / Cfg::EPB;
    const int per_sm = std::max(1, std::min(2048 / Cfg::BT, (int)((size_t)220 * 1024 / (smem + 1024))));
    const int grid = (int)std::min<ptrdiff_t>(nbatches, (ptrdiff_t)sm_count() * per_sm);
    kern<<<grid, Cfg::BT, smem, s>>>(ne, idx, local, x, y);
    B2_CHECK_LAUNCH();
    return B200FEM_SUCCESS;
}

template <typename T, int BS, class Local>
static int launch_micro_level(int level, ptrdiff_t ne, const int32_t *idx, const Local &local, const T *x, T *y, cudaStream_t s) {
    switch (level) {
        case 2: return launch_micro<T, 2, BS, Local>(ne, idx, local, x, y, s);
        case 4: return launch_micro<T, 4, BS, Local>(ne, idx, local, x, y, s);
        case 8: return launch_micro<T, 8, BS, Local>(ne, idx, local, x, y, s);
        default: B2_FAIL("sshex8 operator: level must be 2, 4 or 8");
    }
}

struct Sshex8EM final : OpBase {
    DeviceArray<int32_t> idx_aos;
    DeviceArray<unsigned char> matrices;  // [ne][(8 bs)^2] of real_type

    template <typename T>
    int apply_t(const T *x, T *y, cudaStream_t s) {
        const T *M = reinterpret_cast<const T *>(matrices.ptr);
        if (block_size == 1) return launch_micro_level<T, 1>(level, ne, idx_aos.ptr, MicroMatrixLocal<T, 1>{M}, x, y, s);
        return launch_micro_level<T, 3>(level, ne, idx_aos.ptr, MicroMatrixLocal<T, 3>{M}, x, y, s);
    }
    int apply(const void *x, void *y, bool overwrite, cudaStream_t s) override {
        if (overwrite) B2_CUDA(cudaMemsetAsync(y, 0, (size_t)nn * block_size * real_type, s));
        return real_type == B200FEM_F64 ? apply_t<double>((const double *)x, (double *)y, s) : apply_t<float>((const float *)x, (float *)y, s);
    }
    int diag(void *d, cudaStream_t s) override {
        const ptrdiff_t total = ne * nxe;
        if (total == 0) return B200FEM_SUCCESS;
        const int grid = (int)std::min<ptrdiff_t>((total + 255) / 256, (ptrdiff_t)sm_count() * 16);
        if (real_type == B200FEM_F64) {
            if (block_size == 1)
                sshex8_em_diag_kernel<double, 1><<<grid, 256, 0, s>>>(level, ne, idx_aos.ptr, (const double *)matrices.ptr, (double *)d);
            else
                sshex8_em_diag_kernel<double, 3><<<grid, 256, 0, s>>>(level, ne, idx_aos.ptr, (const double *)matrices.ptr, (double *)d);
        } else {
            if (block_size == 1)
                sshex8_em_diag_kernel<float, 1><<<grid, 256, 0, s>>>(level, ne, idx_aos.ptr, (const float *)matrices.ptr, (float *)d);
            else
                sshex8_em_diag_kernel<float, 3><<<grid, 256, 0, s>>>(level, ne, idx_aos.ptr, (const float *)matrices.ptr, (float *)d);
        }
        B2_CHECK_LAUNCH();
        return B200FEM_SUCCESS;
    }
    int launches(bool overwrite) const override { return 1 + (overwrite ? 1 : 0); }
};

struct Sshex8NonAffineLaplacian final : OpBase {
    DeviceArray<int32_t> idx_aos;
    DeviceArray<float> points;  // [3][nn]
    template <typename T>
    int apply_t(const T *x, T *y, cudaStream_t s) {
        return launch_micro_level<T, 1>(level, ne, idx_aos.ptr, NonAffineLaplacianLocal<T>{idx_aos.ptr, points.ptr, nn, level, nxe}, x, y, s);
    }
    int apply(const void *x, void *y, bool overwrite, cudaStream_t s) override {
        if (overwrite) B2_CUDA(cudaMemsetAsync(y, 0, (size_t)nn * real_type, s));
        return real_type == B200FEM_F64 ? apply_t<double>((const double *)x, (double *)y, s) : apply_t<float>((const float *)x, (float *)y, s);
    }
    int launches(bool overwrite) const override { return 1 + (overwrite ? 1 : 0); }
};

OpBase *make_sshex8_em(int level, ptrdiff_t ne, ptrdiff_t nn, const int32_t *const *elements, int block_size,
                       const void *element_matrix, int real_type) {
    if (level != 2 && level != 4 && level != 8) {
        set_error("sshex8 EM operator: level must be 2, 4 or 8");
        return nullptr;
    }
    if (block_size != 1 && block_size != 3) {
        set_error("sshex8 EM operator: block_size must be 1 (Laplacian) or 3 (linear elasticity)");
        return nullptr;
    }
    auto *op = new Sshex8EM();
    op->elem_type = B200FEM_SSHEX8;
    op->level = level;
    op->real_type = real_type;
    op->block_size = block_size;
    op->ne = ne;
    op->nn = nn;
    op->nxe = (level + 1) * (level + 1) * (level + 1);
    const size_t nb = (size_t)8 * block_size;
    op->geom_bytes_per_elem = double(nb * nb * real_type);  // the stored matrix is this operator's "geometry" stream
    if (upload_aos(level, ne, nn, elements, op->idx_aos) ||
        op->matrices.upload((const unsigned char *)element_matrix, (size_t)ne * nb * nb * real_type)) {
        delete op;
        return nullptr;
    }
    return op;
}

OpBase *make_sshex8_nonaffine_laplacian(int level, ptrdiff_t ne, ptrdiff_t nn, const int32_t *const *elements,
                                        const float *const *points, int real_type) {
    if (level != 2 && level != 4 && level != 8) {
        set_error("sshex8 laplacian (from points): level must be 2, 4 or 8");
        return nullptr;
    }
    auto *op = new Sshex8NonAffineLaplacian();
    op->elem_type = B200FEM_SSHEX8;
    op->level = level;
    op->real_type = real_type;
    op->ne = ne;
    op->nn = nn;
    op->nxe = (level + 1) * (level + 1) * (level + 1);
    op->geom_bytes_per_elem = 8 * 3 * 4;  // the 8 corner coordinates
    bool bad = upload_aos(level, ne, nn, elements, op->idx_aos) != 0 || op->points.alloc((size_t)3 * nn) != 0;
    for (int d = 0; d < 3 && !bad; ++d)
        bad = cudaMemcpy(op->points.ptr + (size_t)d * nn, points[d], (size_t)nn * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess;
    if (bad) {
        delete op;
        return nullptr;
    }
    return op;
}

template <typename T>
static int em_build_laplacian(int level, ptrdiff_t ne, const float *fff, void *values_host) {
    DeviceArray<float> d_fff;
    DeviceArray<T> d_out;
    if (d_fff.upload(fff, (size_t)ne * 6) || d_out.alloc((size_t)ne * 64)) return B200FEM_FAILURE;
    if (ne) {
        sshex8_laplacian_em_kernel<T><<<(int)std::min<ptrdiff_t>((ne + 127) / 128, 4096), 128>>>(level, ne, d_fff.ptr, d_out.ptr);
        B2_CHECK_LAUNCH();
        B2_CUDA(cudaMemcpy(values_host, d_out.ptr, (size_t)ne * 64 * sizeof(T), cudaMemcpyDeviceToHost));
    }
    return B200FEM_SUCCESS;
}

template <typename T>
static int em_build_le(int level, ptrdiff_t ne, const float *adjugate, const float *determinant, double mu, double lambda, void *values_host) {
    DeviceArray<float> d_adj, d_det;
    DeviceArray<T> d_out;
    if (d_adj.upload(adjugate, (size_t)ne * 9) || d_det.upload(determinant, (size_t)ne) || d_out.alloc((size_t)ne * 576)) return B200FEM_FAILURE;
    if (ne) {
        sshex8_le_em_kernel<T><<<(int)std::min<ptrdiff_t>((ne + 63) / 64, 4096), 64>>>(level, ne, d_adj.ptr, d_det.ptr, (T)mu, (T)lambda, d_out.ptr);
        B2_CHECK_LAUNCH();
        B2_CUDA(cudaMemcpy(values_host, d_out.ptr, (size_t)ne * 576 * sizeof(T), cudaMemcpyDeviceToHost));
    }
    return B200FEM_SUCCESS;
}

int sshex8_laplacian_element_matrix_build(int level, ptrdiff_t ne, const float *fff, int real_type, void *values_host) {
    return real_type == B200FEM_F64 ? em_build_laplacian<double>(level, ne, fff, values_host) : em_build_laplacian<float>(level, ne, fff, values_host);
}
int sshex8_le_element_matrix_build(int level, ptrdiff_t ne, const float *adjugate, const float *determinant, double mu, double lambda,
                                   int real_type, void *values_host) {
    return real_type == B200FEM_F64 ? em_build_le<double>(level, ne, adjugate, determinant, mu, lambda, values_host)
                                    : em_build_le<float>(level, ne, adjugate, determinant, mu, lambda, values_host);
}

}  // namespace b200fem
