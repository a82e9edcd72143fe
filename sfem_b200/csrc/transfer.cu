// Multigrid transfer operators between two lattices of the same semi-structured macro mesh, and the
// block-Jacobi smoother kernels.
//
// Replaces sshex8_restrict / sshex8_prolongate / sshex8_hierarchical_restriction /
// sshex8_hierarchical_prolongation / sshex8_element_node_incidence_count
// (operators/sshex8/sshex8_interpolate.cpp:12-127, 129-237, 239-257, 259-615, 633-775 of the reference) and
// the ShiftableBlockSymJacobi kernels (algebra/openmp/sfem_openmp_ShiftableJacobi.hpp:30-226).
//
// Both transfers are tensor products of the 1-D hat-function matrix W[c][f] = max(0, 1 - |f / step - c|)
// (step = fine_level / coarse_level): prolongation = W^T (x) W^T (x) W^T, restriction = its transpose applied
// to the fine values divided by their element-incidence count. One CTA works on one macro element at a time:
// the lattice goes through shared memory, one 1-D pass per direction (sum factorisation: 3 short passes
// instead of an (Lf+1)^3 x (Lc+1)^3 dense product), coalesced gathers through the AoS connectivity,
// RED.ADD (restriction) / plain stores (prolongation) on the way out.
#include <algorithm>
#include <vector>

#include "common.cuh"

namespace b200fem {

struct Transfer {
    ptrdiff_t ne = 0, n_fine = 0, n_coarse = 0;
    int lf = 0, lc = 0;  // fine / coarse lattice levels
    DeviceArray<int32_t> fine_idx;    // [ne][(lf+1)^3]
    DeviceArray<int32_t> coarse_idx;  // [ne][(lc+1)^3]
    DeviceArray<uint16_t> count;      // [n_fine] element incidence of every fine node
};

__device__ __forceinline__ double hat_weight(int f, int c, int step) {
    const int d = f - c * step;
    const int a = d < 0 ? -d : d;
    return a >= step ? 0.0 : 1.0 - (double)a / (double)step;
}

// RESTRICT = true : fine -> coarse (to += W from/count), false: coarse -> fine (to = W^T from)
template <typename T, bool RESTRICT>
__global__ void __launch_bounds__(128)
    sshex8_transfer_kernel(const ptrdiff_t ne, const int lf, const int lc, const int32_t *__restrict__ fine_idx,
                           const int32_t *__restrict__ coarse_idx, const uint16_t *__restrict__ count, const int vec_size,
                           const T *__restrict__ from, T *__restrict__ to) {
    const int nf = lf + 1, nc = lc + 1, step = lf / lc;
    const int nf3 = nf * nf * nf, nc3 = nc * nc * nc;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *a = reinterpret_cast<T *>(smem_raw);  // up to nf^3
    T *b = a + nf3;                          // up to nf^2 * nc
    const int tid = threadIdx.x, nt = blockDim.x;

    for (ptrdiff_t e = blockIdx.x; e < ne; e += gridDim.x) {
        const int32_t *fi = fine_idx + e * nf3;
        const int32_t *ci = coarse_idx + e * nc3;
        for (int d = 0; d < vec_size; ++d) {
            if (RESTRICT) {
                for (int n = tid; n < nf3; n += nt) {
                    const ptrdiff_t id = fi[n];
                    a[n] = from[id * vec_size + d] / (T)count[id];
                }
                __syncthreads();
                // x: b[z][y][cx] = sum_x W[cx][x] a[z][y][x]
                for (int o = tid; o < nf * nf * nc; o += nt) {
                    const int cx = o % nc, zy = o / nc;
                    T s = 0;
                    const int x0 = max(0, (cx - 1) * step + 1), x1 = min(lf, (cx + 1) * step - 1);
                    for (int x = x0; x <= x1; ++x) s += (T)hat_weight(x, cx, step) * a[zy * nf + x];
                    b[o] = s;
                }
                __syncthreads();
                // y: a[z][cy][cx] = sum_y W[cy][y] b[z][y][cx]
                for (int o = tid; o < nf * nc * nc; o += nt) {
                    const int cx = o % nc, cy = (o / nc) % nc, z = o / (nc * nc);
                    T s = 0;
                    const int y0 = max(0, (cy - 1) * step + 1), y1 = min(lf, (cy + 1) * step - 1);
                    for (int y = y0; y <= y1; ++y) s += (T)hat_weight(y, cy, step) * b[(z * nf + y) * nc + cx];
                    a[o] = s;
                }
                __syncthreads();
                // z: out[cz][cy][cx] = sum_z W[cz][z] a[z][cy][cx]
                for (int o = tid; o < nc3; o += nt) {
                    const int cz = o / (nc * nc), r = o - cz * nc * nc;
                    T s = 0;
                    const int z0 = max(0, (cz - 1) * step + 1), z1 = min(lf, (cz + 1) * step - 1);
                    for (int z = z0; z <= z1; ++z) s += (T)hat_weight(z, cz, step) * a[z * nc * nc + r];
                    red_add(&to[(ptrdiff_t)ci[o] * vec_size + d], s);
                }
                __syncthreads();
            } else {
                for (int n = tid; n < nc3; n += nt) a[n] = from[(ptrdiff_t)ci[n] * vec_size + d];
                __syncthreads();
                // z: b[z][cy][cx] = sum_cz W[cz][z] a[cz][cy][cx]
                for (int o = tid; o < nf * nc * nc; o += nt) {
                    const int z = o / (nc * nc), r = o - z * nc * nc;
                    const int c0 = min(z / step, lc - 1);
                    b[o] = (T)hat_weight(z, c0, step) * a[c0 * nc * nc + r] + (T)hat_weight(z, c0 + 1, step) * a[(c0 + 1) * nc * nc + r];
                }
                __syncthreads();
                // y: a[z][y][cx]
                for (int o = tid; o < nf * nf * nc; o += nt) {
                    const int cx = o % nc, y = (o / nc) % nf, z = o / (nc * nf);
                    const int c0 = min(y / step, lc - 1);
                    a[o] = (T)hat_weight(y, c0, step) * b[(z * nc + c0) * nc + cx] + (T)hat_weight(y, c0 + 1, step) * b[(z * nc + c0 + 1) * nc + cx];
                }
                __syncthreads();
                // x: out[z][y][x]
                for (int n = tid; n < nf3; n += nt) {
                    const int x = n % nf, zy = n / nf;
                    const int c0 = min(x / step, lc - 1);
                    const T v = (T)hat_weight(x, c0, step) * a[zy * nc + c0] + (T)hat_weight(x, c0 + 1, step) * a[zy * nc + c0 + 1];
                    to[(ptrdiff_t)fi[n] * vec_size + d] = v;
                }
                __syncthreads();
            }
        }
    }
}

// ---- block-symmetric Jacobi (3x3 blocks)
template <typename T>
__global__ void __launch_bounds__(256)
    block_sym_jacobi_setup_kernel(const ptrdiff_t n_blocks, const T *__restrict__ sym6, const T *__restrict__ shift6,
                                  const unsigned char *__restrict__ mask, const T relaxation, T *__restrict__ inv9) {
    for (ptrdiff_t b = (ptrdiff_t)blockIdx.x * blockDim.x + threadIdx.x; b < n_blocks; b += (ptrdiff_t)gridDim.x * blockDim.x) {
        T d[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) d[k] = sym6[b * 6 + k] + (shift6 ? shift6[b * 6 + k] : T(0));
        T m[9] = {d[0], d[1], d[2], d[1], d[3], d[4], d[2], d[4], d[5]};
        if (mask) {
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                const ptrdiff_t i = b * 3 + r;
                if ((mask[i >> 3] >> (i & 7)) & 1) {
                    m[r * 3 + 0] = m[r * 3 + 1] = m[r * 3 + 2] = T(0);
                    m[r * 3 + r] = T(1);
                }
            }
        }
        // inverse3 of algebra/openmp/sfem_openmp_ShiftableJacobi.hpp:37-71
        const T x0 = m[4] * m[8], x1 = m[5] * m[7], x2 = m[1] * m[5], x3 = m[1] * m[8], x4 = m[2] * m[4];
        const T x5 = T(1) / (m[0] * x0 - m[0] * x1 + m[2] * m[3] * m[7] - m[3] * x3 + m[6] * x2 - m[6] * x4);
        T *o = inv9 + b * 9;
        o[0] = relaxation * (x5 * (x0 - x1));
        o[1] = relaxation * (x5 * (m[2] * m[7] - x3));
        o[2] = relaxation * (x5 * (x2 - x4));
        o[3] = relaxation * (x5 * (-m[3] * m[8] + m[5] * m[6]));
        o[4] = relaxation * (x5 * (m[0] * m[8] - m[2] * m[6]));
        o[5] = relaxation * (x5 * (-m[0] * m[5] + m[2] * m[3]));
        o[6] = relaxation * (x5 * (m[3] * m[7] - m[4] * m[6]));
        o[7] = relaxation * (x5 * (-m[0] * m[7] + m[1] * m[6]));
        o[8] = relaxation * (x5 * (m[0] * m[4] - m[1] * m[3]));
    }
}

template <typename T>
__global__ void __launch_bounds__(256)
    block_jacobi_apply_kernel(const ptrdiff_t n_blocks, const T *__restrict__ inv9, const T *__restrict__ x,
                              T *__restrict__ y) {
    for (ptrdiff_t b = (ptrdiff_t)blockIdx.x * blockDim.x + threadIdx.x; b < n_blocks; b += (ptrdiff_t)gridDim.x * blockDim.x) {
        const T *d = inv9 + b * 9;
        const T x0 = x[b * 3 + 0], x1 = x[b * 3 + 1], x2 = x[b * 3 + 2];
        y[b * 3 + 0] += d[0] * x0 + d[1] * x1 + d[2] * x2;
        y[b * 3 + 1] += d[3] * x0 + d[4] * x1 + d[5] * x2;
        y[b * 3 + 2] += d[6] * x0 + d[7] * x1 + d[8] * x2;
    }
}

// MixedPrecisionShiftableBlockSymJacobi<HP, LP> (algebra/sfem_MixedPrecisionShiftableBlockSymJacobi.hpp): the block
// diagonal arrives in HP, is converted to LP (sym_diag_to_diag), masked, inverted and scaled in LP
// (algebra/openmp/sfem_openmp_ShiftableJacobi.hpp:31-98, 227-243), and the apply mixes LP blocks with HP vectors.
template <typename HP, typename LP>
__global__ void __launch_bounds__(256)
    block_sym_jacobi_setup_mixed_kernel(const ptrdiff_t n_blocks, const HP *__restrict__ sym6, const HP *__restrict__ shift6,
                                        const unsigned char *__restrict__ mask, const LP relaxation, LP *__restrict__ inv9) {
    for (ptrdiff_t b = (ptrdiff_t)blockIdx.x * blockDim.x + threadIdx.x; b < n_blocks; b += (ptrdiff_t)gridDim.x * blockDim.x) {
        LP d[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            d[k] = (LP)sym6[b * 6 + k];                 // sym_diag_to_diag<HP, LP>
            if (shift6) d[k] += (LP)shift6[b * 6 + k];  // add_sym_diag_to_diag<HP, LP>
        }
        LP m[9] = {d[0], d[1], d[2], d[1], d[3], d[4], d[2], d[4], d[5]};
        if (mask) {
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                const ptrdiff_t i = b * 3 + r;
                if ((mask[i >> 3] >> (i & 7)) & 1) {
                    m[r * 3 + 0] = m[r * 3 + 1] = m[r * 3 + 2] = LP(0);
                    m[r * 3 + r] = LP(1);
                }
            }
        }
        const LP x0 = m[4] * m[8], x1 = m[5] * m[7], x2 = m[1] * m[5], x3 = m[1] * m[8], x4 = m[2] * m[4];
        const LP x5 = LP(1) / (m[0] * x0 - m[0] * x1 + m[2] * m[3] * m[7] - m[3] * x3 + m[6] * x2 - m[6] * x4);
        LP *o = inv9 + b * 9;
        o[0] = relaxation * (x5 * (x0 - x1));
        o[1] = relaxation * (x5 * (m[2] * m[7] - x3));
        o[2] = relaxation * (x5 * (x2 - x4));
        o[3] = relaxation * (x5 * (-m[3] * m[8] + m[5] * m[6]));
        o[4] = relaxation * (x5 * (m[0] * m[8] - m[2] * m[6]));
        o[5] = relaxation * (x5 * (-m[0] * m[5] + m[2] * m[3]));
        o[6] = relaxation * (x5 * (m[3] * m[7] - m[4] * m[6]));
        o[7] = relaxation * (x5 * (-m[0] * m[7] + m[1] * m[6]));
        o[8] = relaxation * (x5 * (m[0] * m[4] - m[1] * m[3]));
    }
}

template <typename HP, typename LP>
__global__ void __launch_bounds__(256)
    block_jacobi_apply_mixed_kernel(const ptrdiff_t n_blocks, const LP *__restrict__ inv9, const HP *__restrict__ x, HP *__restrict__ y) {
    for (ptrdiff_t b = (ptrdiff_t)blockIdx.x * blockDim.x + threadIdx.x; b < n_blocks; b += (ptrdiff_t)gridDim.x * blockDim.x) {
        const LP *d = inv9 + b * 9;
        const HP x0 = x[b * 3 + 0], x1 = x[b * 3 + 1], x2 = x[b * 3 + 2];
        y[b * 3 + 0] += (HP)d[0] * x0 + (HP)d[1] * x1 + (HP)d[2] * x2;
        y[b * 3 + 1] += (HP)d[3] * x0 + (HP)d[4] * x1 + (HP)d[5] * x2;
        y[b * 3 + 2] += (HP)d[6] * x0 + (HP)d[7] * x1 + (HP)d[8] * x2;
    }
}

}  // namespace b200fem

using namespace b200fem;

struct b200fem_transfer_s {
    Transfer t;
};

extern "C" {

int b200fem_sshex8_transfer_create(b200fem_transfer_t *out, ptrdiff_t ne, int fine_level, int fine_stride,
                                   const int32_t *const *fine_elements, ptrdiff_t n_fine_nodes, int coarse_level,
                                   int coarse_stride, const int32_t *const *coarse_elements, ptrdiff_t n_coarse_nodes) {
    if (!out) B2_FAIL("transfer output pointer is NULL");
    *out = nullptr;
    if (ne < 0 || n_fine_nodes < 0 || n_coarse_nodes < 0) B2_FAIL("negative sizes");
    if (fine_level < 1 || coarse_level < 1 || fine_level % coarse_level != 0 || fine_level == coarse_level)
        B2_FAIL("nested meshes requirement: fine_level must be a proper multiple of coarse_level");
    if (fine_level > 16) B2_FAIL("sshex8 transfer: fine_level > 16 not supported");
    if (fine_stride < 1 || coarse_stride < 1) B2_FAIL("strides must be >= 1");
    if (ne > 0 && (!fine_elements || !coarse_elements)) B2_FAIL("elements is NULL");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) B2_FAIL("no CUDA device: b200fem has no CPU fallback");

    auto *h = new b200fem_transfer_s();
    Transfer &t = h->t;
    t.ne = ne; t.n_fine = n_fine_nodes; t.n_coarse = n_coarse_nodes; t.lf = fine_level; t.lc = coarse_level;
    const int nf = fine_level + 1, nc = coarse_level + 1;
    const int nf3 = nf * nf * nf, nc3 = nc * nc * nc;
    // strided views: lattice node (x,y,z) of the level-l lattice is row lidx(l*stride, x*stride, ...) of `elements`
    auto lidx = [](int L, int x, int y, int z) { return z * (L + 1) * (L + 1) + y * (L + 1) + x; };
    std::vector<int32_t> fi((size_t)ne * nf3), ci((size_t)ne * nc3);
    std::vector<uint16_t> cnt((size_t)n_fine_nodes, 0);
    bool ok = true;
    for (int z = 0; z < nf && ok; ++z)
        for (int y = 0; y < nf && ok; ++y)
            for (int x = 0; x < nf; ++x) {
                const int32_t *row = fine_elements[lidx(fine_level * fine_stride, x * fine_stride, y * fine_stride, z * fine_stride)];
                const int n = lidx(fine_level, x, y, z);
                for (ptrdiff_t e = 0; e < ne; ++e) {
                    const int32_t id = row[e];
                    if (id < 0 || id >= n_fine_nodes) { ok = false; break; }
                    fi[(size_t)e * nf3 + n] = id;
                    cnt[(size_t)id]++;  // sshex8_element_node_incidence_count (sshex8_interpolate.cpp:239-257)
                }
            }
    for (int z = 0; z < nc && ok; ++z)
        for (int y = 0; y < nc && ok; ++y)
            for (int x = 0; x < nc; ++x) {
                const int32_t *row = coarse_elements[lidx(coarse_level * coarse_stride, x * coarse_stride, y * coarse_stride, z * coarse_stride)];
                const int n = lidx(coarse_level, x, y, z);
                for (ptrdiff_t e = 0; e < ne; ++e) {
                    const int32_t id = row[e];
                    if (id < 0 || id >= n_coarse_nodes) { ok = false; break; }
                    ci[(size_t)e * nc3 + n] = id;
                }
            }
    if (!ok) {
        delete h;
        B2_FAIL("sshex8 transfer: node index out of range");
    }
    if (t.fine_idx.upload(fi.data(), fi.size()) || t.coarse_idx.upload(ci.data(), ci.size()) ||
        t.count.upload(cnt.data(), cnt.size())) {
        delete h;
        return B200FEM_FAILURE;
    }
    *out = h;
    return B200FEM_SUCCESS;
}

int b200fem_transfer_destroy(b200fem_transfer_t t) {
    delete t;
    return B200FEM_SUCCESS;
}

ptrdiff_t b200fem_transfer_fine_nodes(b200fem_transfer_t t) { return t ? t->t.n_fine : -1; }
ptrdiff_t b200fem_transfer_coarse_nodes(b200fem_transfer_t t) { return t ? t->t.n_coarse : -1; }

int b200fem_transfer_incidence_count(b200fem_transfer_t t, uint16_t *count_host) {
    if (!t || !count_host) B2_FAIL("invalid transfer handle");
    if (t->t.count.n) B2_CUDA(cudaMemcpy(count_host, t->t.count.ptr, t->t.count.bytes(), cudaMemcpyDeviceToHost));
    return B200FEM_SUCCESS;
}

int b200fem_transfer_set_incidence_count(b200fem_transfer_t t, const uint16_t *count_host) {
    if (!t || !count_host) B2_FAIL("invalid transfer handle");
    if (t->t.count.n) B2_CUDA(cudaMemcpy(t->t.count.ptr, count_host, t->t.count.bytes(), cudaMemcpyHostToDevice));
    return B200FEM_SUCCESS;
}

}  // extern "C"

template <bool RESTRICT>
static int transfer_launch(b200fem_transfer_t h, int real_type, int vec_size, const void *from, void *to, void *stream) {
    if (!h) B2_FAIL("invalid transfer handle");
    if (real_type != B200FEM_F32 && real_type != B200FEM_F64) B2_FAIL("real_type must be B200FEM_F32 or B200FEM_F64");
    if (vec_size < 1) B2_FAIL("vec_size must be >= 1");
    const Transfer &t = h->t;
    if (t.ne == 0) return B200FEM_SUCCESS;
    if (!from || !to) B2_FAIL("transfer: NULL vector");
    const int nf = t.lf + 1, nc = t.lc + 1;
    const size_t smem = (size_t)(nf * nf * nf + nf * nf * nc) * real_type;
    const int grid = (int)std::min<ptrdiff_t>(t.ne, (ptrdiff_t)sm_count() * 8);
    cudaStream_t s = as_stream(stream);
    if (real_type == B200FEM_F64) {
        auto k = sshex8_transfer_kernel<double, RESTRICT>;
        if (smem > 48 * 1024) B2_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k<<<grid, 128, smem, s>>>(t.ne, t.lf, t.lc, t.fine_idx.ptr, t.coarse_idx.ptr, t.count.ptr, vec_size, (const double *)from, (double *)to);
    } else {
        auto k = sshex8_transfer_kernel<float, RESTRICT>;
        if (smem > 48 * 1024) B2_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k<<<grid, 128, smem, s>>>(t.ne, t.lf, t.lc, t.fine_idx.ptr, t.coarse_idx.ptr, t.count.ptr, vec_size, (const float *)from, (float *)to);
    }
    B2_CHECK_LAUNCH();
    return B200FEM_SUCCESS;
}

extern "C" {

int b200fem_transfer_restrict(b200fem_transfer_t t, int real_type, int vec_size, const void *from, void *to, void *stream) {
    return transfer_launch<true>(t, real_type, vec_size, from, to, stream);
}

int b200fem_transfer_prolongate(b200fem_transfer_t t, int real_type, int vec_size, const void *from, void *to, void *stream) {
    return transfer_launch<false>(t, real_type, vec_size, from, to, stream);
}

int b200fem_block_sym_jacobi_setup(int real_type, ptrdiff_t n_blocks, const void *sym_diag6, const void *shift6,
                                   const void *constraints_mask, double relaxation, void *inv_diag9, void *stream) {
    if (n_blocks < 0) B2_FAIL("negative size");
    if (n_blocks == 0) return B200FEM_SUCCESS;
    if (!sym_diag6 || !inv_diag9) B2_FAIL("block_sym_jacobi_setup: NULL array");
    const int grid = (int)std::min<ptrdiff_t>((n_blocks + 255) / 256, (ptrdiff_t)sm_count() * 8);
    if (real_type == B200FEM_F64)
        block_sym_jacobi_setup_kernel<double><<<grid, 256, 0, as_stream(stream)>>>(
            n_blocks, (const double *)sym_diag6, (const double *)shift6, (const unsigned char *)constraints_mask, relaxation, (double *)inv_diag9);
    else if (real_type == B200FEM_F32)
        block_sym_jacobi_setup_kernel<float><<<grid, 256, 0, as_stream(stream)>>>(
            n_blocks, (const float *)sym_diag6, (const float *)shift6, (const unsigned char *)constraints_mask, (float)relaxation, (float *)inv_diag9);
    else
        B2_FAIL("real_type must be B200FEM_F32 or B200FEM_F64");
    B2_CHECK_LAUNCH();
    return B200FEM_SUCCESS;
}

int b200fem_block_sym_jacobi_setup_mixed(ptrdiff_t n_blocks, const double *sym_diag6, const double *shift6, const void *constraints_mask,
                                         double relaxation, float *inv_diag9, void *stream) {
    if (n_blocks < 0) B2_FAIL("negative size");
    if (n_blocks == 0) return B200FEM_SUCCESS;
    if (!sym_diag6 || !inv_diag9) B2_FAIL("block_sym_jacobi_setup_mixed: NULL array");
    const int grid = (int)std::min<ptrdiff_t>((n_blocks + 255) / 256, (ptrdiff_t)sm_count() * 8);
    block_sym_jacobi_setup_mixed_kernel<double, float><<<grid, 256, 0, as_stream(stream)>>>(
        n_blocks, sym_diag6, shift6, (const unsigned char *)constraints_mask, (float)relaxation, inv_diag9);
    B2_CHECK_LAUNCH();
    return B200FEM_SUCCESS;
}

int b200fem_block_jacobi_apply_mixed(ptrdiff_t n_blocks, const float *inv_diag9, const double *x, double *y, void *stream) {
    if (n_blocks < 0) B2_FAIL("negative size");
    if (n_blocks == 0) return B200FEM_SUCCESS;
    if (!inv_diag9 || !x || !y) B2_FAIL("block_jacobi_apply_mixed: NULL array");
    const int grid = (int)std::min<ptrdiff_t>((n_blocks + 255) / 256, (ptrdiff_t)sm_count() * 8);
    block_jacobi_apply_mixed_kernel<double, float><<<grid, 256, 0, as_stream(stream)>>>(n_blocks, inv_diag9, x, y);
    B2_CHECK_LAUNCH();
    return B200FEM_SUCCESS;
}

int b200fem_block_jacobi_apply(int real_type, ptrdiff_t n_blocks, const void *inv_diag9, const void *x, void *y, void *stream) {
    if (n_blocks < 0) B2_FAIL("negative size");
    if (n_blocks == 0) return B200FEM_SUCCESS;
    if (!inv_diag9 || !x || !y) B2_FAIL("block_jacobi_apply: NULL array");
    const int grid = (int)std::min<ptrdiff_t>((n_blocks + 255) / 256, (ptrdiff_t)sm_count() * 8);
    if (real_type == B200FEM_F64)
        block_jacobi_apply_kernel<double><<<grid, 256, 0, as_stream(stream)>>>(n_blocks, (const double *)inv_diag9, (const double *)x, (double *)y);
    else if (real_type == B200FEM_F32)
        block_jacobi_apply_kernel<float><<<grid, 256, 0, as_stream(stream)>>>(n_blocks, (const float *)inv_diag9, (const float *)x, (float *)y);
    else
        B2_FAIL("real_type must be B200FEM_F32 or B200FEM_F64");
    B2_CHECK_LAUNCH();
    return B200FEM_SUCCESS;
}

}  // extern "C"
