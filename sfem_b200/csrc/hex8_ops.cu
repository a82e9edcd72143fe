// hex8 (affine trilinear) Laplacian and linear elasticity, matrix-free, one thread per element.
//
// Laplacian replaces hex8_laplacian_apply_opt (operators/hex8/hex8_laplacian.cpp:309-345 of the
// reference; local op hex8_laplacian_apply_fff_integral, hex8_laplacian_inline_cpu.hpp:332-443).
// Linear elasticity replaces affine_hex8_linear_elasticity_apply_adjugate_soa
// (operators/hex8/hex8_linear_elasticity.cpp:625-787: 2x2x2 Gauss, sum-factorised, node order
// permuted VTK -> lexicographic :43-53) and affine_hex8_linear_elasticity_diag /
// _block_diag_sym (:1043-1200).
//
// Both local operators are evaluated in the tensor sum/difference basis of hex8_local.cuh. For an
// affine element the 2x2x2 Gauss rule of the reference integrates the (per-direction quadratic)
// integrands exactly, so the closed-form reference matrices G_jl = int d_j phi_m d_l phi_n give the
// same element vector:
//     out_a = sum_b sum_{j,l} T^{ab}_{jl} G_jl u_b,
//     T^{ab}_{jl} = [ lambda A_ja A_lb + mu delta_ab (A A^T)_jl + mu A_jb A_la ] / det ,  A = adjugate
// (A/det = J^-1 as used at linear_elasticity_d3_tensor_product_local.hpp:264-292: grad_x u_a[k] =
// sum_j d_j u_a A[j*3+k] / det). ~450 flops per element instead of ~2500.
//
// HBM layout: connectivity SoA int32 [8][ne], FFF SoA fp32 [6][ne], adjugate SoA fp32 [9][ne] (the
// layout the reference Op caches for HEX8, frontend/ops/sfem_LinearElasticity.cpp:275-277),
// determinant fp32 [ne]; all coalesced over elements and streamed once. Vectors: scalar [nn] or AoS
// [nn*3]; gathers/scatters go through L2.
#include <string.h>

#include <vector>

#include "coloring.cuh"
#include "common.cuh"
#include "hex8_le_core.cuh"
#include "hex8_local.cuh"
#include "patch.cuh"

namespace b200fem {

template <typename T, bool COLORED>
__global__ void __launch_bounds__(256)
    hex8_laplacian_apply_kernel(const ptrdiff_t begin, const ptrdiff_t end, const ptrdiff_t ne,
                                const int32_t *__restrict__ elements, const float *__restrict__ fff,
                                const int32_t *__restrict__ perm, const T *__restrict__ u, T *__restrict__ y) {
    for (ptrdiff_t i = begin + (ptrdiff_t)blockIdx.x * blockDim.x + threadIdx.x; i < end;
         i += (ptrdiff_t)gridDim.x * blockDim.x) {
        const ptrdiff_t e = COLORED ? (ptrdiff_t)perm[i] : i;
        int32_t ev[8];
        T ul[8], f[6], ol[8];
#pragma unroll
        for (int v = 0; v < 8; ++v) ev[v] = ldg_stream(&elements[v * ne + e]);
#pragma unroll
        for (int k = 0; k < 6; ++k) f[k] = (T)ldg_stream(&fff[k * ne + e]);
#pragma unroll
        for (int v = 0; v < 8; ++v) ul[kVtkToLex[v]] = u[ev[v]];
        hex8_laplacian_local(f, ul, ol);
#pragma unroll
        for (int v = 0; v < 8; ++v) {
            if (COLORED)
                y[ev[v]] += ol[kVtkToLex[v]];
            else
                red_add(&y[ev[v]], ol[kVtkToLex[v]]);
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(256)
    hex8_laplacian_diag_kernel(const ptrdiff_t ne, const int32_t *__restrict__ elements,
                               const float *__restrict__ fff, T *__restrict__ diag) {
    for (ptrdiff_t e = (ptrdiff_t)blockIdx.x * blockDim.x + threadIdx.x; e < ne; e += (ptrdiff_t)gridDim.x * blockDim.x) {
        T f[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) f[k] = (T)fff[k * ne + e];
#pragma unroll
        for (int v = 0; v < 8; ++v) red_add(&diag[elements[v * ne + e]], hex8_laplacian_diag_entry(f, kVtkToLex[v]));
    }
}

// VectorLaplacian on hex8: the scalar operator on every component of an AoS field [nn * bs]
// (affine_hex8_vector_laplacian_apply_fff, operators/hex8/hex8_vector_laplacian.cpp:78-138: one element matrix, applied per component)
template <typename T, bool DIAG>
__global__ void __launch_bounds__(256)
    hex8_vector_laplacian_kernel(const ptrdiff_t ne, const int bs, const int32_t *__restrict__ elements, const float *__restrict__ fff,
                                 const T *__restrict__ u, T *__restrict__ y) {
    for (ptrdiff_t e = (ptrdiff_t)blockIdx.x * blockDim.x + threadIdx.x; e < ne; e += (ptrdiff_t)gridDim.x * blockDim.x) {
        int32_t ev[8];
        T f[6];
#pragma unroll
        for (int v = 0; v < 8; ++v) ev[v] = ldg_stream(&elements[v * ne + e]);
#pragma unroll
        for (int k = 0; k < 6; ++k) f[k] = (T)ldg_stream(&fff[k * ne + e]);
#pragma unroll 1
        for (int d = 0; d < bs; ++d) {
            if (DIAG) {
#pragma unroll
                for (int v = 0; v < 8; ++v) red_add(&y[(ptrdiff_t)ev[v] * bs + d], hex8_laplacian_diag_entry(f, kVtkToLex[v]));
            } else {
                T ul[8], ol[8];
#pragma unroll
                for (int v = 0; v < 8; ++v) ul[kVtkToLex[v]] = u[(ptrdiff_t)ev[v] * bs + d];
                hex8_laplacian_local(f, ul, ol);
#pragma unroll
                for (int v = 0; v < 8; ++v) red_add(&y[(ptrdiff_t)ev[v] * bs + d], ol[kVtkToLex[v]]);
            }
        }
    }
}

// T^{ab} (3x3 in j,l) for one (a,b) pair
template <typename T>
__device__ __forceinline__ void le_coeff(const T A[9], const T AAt_mu[6], const T lam_idet, const T mu_idet, int a,
                                         int b, T t[3][3]) {
    // AAt_mu holds mu * (A A^T) / det, symmetric (00 01 02 11 12 22)
    const int sym[3][3] = {{0, 1, 2}, {1, 3, 4}, {2, 4, 5}};
#pragma unroll
    for (int j = 0; j < 3; ++j) {
#pragma unroll
        for (int l = 0; l < 3; ++l) {
            T v = lam_idet * A[j * 3 + a] * A[l * 3 + b] + mu_idet * A[j * 3 + b] * A[l * 3 + a];
            if (a == b) v += AAt_mu[sym[j][l]];
            t[j][l] = v;
        }
    }
}

template <typename T>
__device__ __forceinline__ void le_prepare(const T A[9], const T det, const T mu, const T lambda, T AAt_mu[6],
                                           T &lam_idet, T &mu_idet) {
    const T idet = T(1) / det;
    lam_idet = lambda * idet;
    mu_idet = mu * idet;
    int k = 0;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
#pragma unroll
        for (int l = j; l < 3; ++l) {
            AAt_mu[k++] = mu_idet * (A[j * 3 + 0] * A[l * 3 + 0] + A[j * 3 + 1] * A[l * 3 + 1] + A[j * 3 + 2] * A[l * 3 + 2]);
        }
    }
}

template <typename T, bool COLORED>
__global__ void __launch_bounds__(128)
    hex8_le_apply_kernel(const ptrdiff_t begin, const ptrdiff_t end, const ptrdiff_t ne,
                         const int32_t *__restrict__ elements, const float *__restrict__ adjugate,
                         const float *__restrict__ determinant, const T mu, const T lambda,
                         const int32_t *__restrict__ perm, const T *__restrict__ u, T *__restrict__ y) {
    for (ptrdiff_t i = begin + (ptrdiff_t)blockIdx.x * blockDim.x + threadIdx.x; i < end;
         i += (ptrdiff_t)gridDim.x * blockDim.x) {
        const ptrdiff_t e = COLORED ? (ptrdiff_t)perm[i] : i;
        int32_t ev[8];
#pragma unroll
        for (int v = 0; v < 8; ++v) ev[v] = ldg_stream(&elements[v * ne + e]);
        T A[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) A[k] = (T)ldg_stream(&adjugate[k * ne + e]);
        const T det = (T)ldg_stream(&determinant[e]);

        // element vector from the gradient-moment form of hex8_le_core.cuh (h = 1: the element is its own macro element):
        // 444 fp64 operations against the ~840 of the per-(a, b) coefficient form this kernel started with
        T c[3][8], O[3][8];
#pragma unroll
        for (int v = 0; v < 8; ++v) {
            const T *p = &u[(ptrdiff_t)ev[v] * 3];
#pragma unroll
            for (int a = 0; a < 3; ++a) c[a][kVtkToLex[v]] = p[a];
        }
        Hex8LEConsts<T> K;
        K.init(A, det, T(1), mu, lambda);
        hex8_le_local_apply(K, c, O);
#pragma unroll
        for (int v = 0; v < 8; ++v) {
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                T *dst = &y[(ptrdiff_t)ev[v] * 3 + a];
                const T val = O[a][kVtkToLex[v]];
                if (COLORED)
                    *dst += val;
                else
                    red_add(dst, val);
            }
        }
    }
}

// diagonal / symmetric 3x3 block diagonal: entry (a,b) at corner `lex` = sum_jl T^{ab}_{jl} G_jl[m][m]
// with G_jj[m][m] = 1/9 and G_jl[m][m] = s_j s_l / 12 (j != l), s = +-1 by corner side.
template <typename T, bool BLOCK>
__global__ void __launch_bounds__(128)
    hex8_le_diag_kernel(const ptrdiff_t ne, const int32_t *__restrict__ elements, const float *__restrict__ adjugate,
                        const float *__restrict__ determinant, const T mu, const T lambda, T *__restrict__ out) {
    for (ptrdiff_t e = (ptrdiff_t)blockIdx.x * blockDim.x + threadIdx.x; e < ne; e += (ptrdiff_t)gridDim.x * blockDim.x) {
        T A[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) A[k] = (T)adjugate[k * ne + e];
        const T det = (T)determinant[e];
        T AAt_mu[6], lam_idet, mu_idet;
        le_prepare(A, det, mu, lambda, AAt_mu, lam_idet, mu_idet);
        int32_t ev[8];
#pragma unroll
        for (int v = 0; v < 8; ++v) ev[v] = elements[v * ne + e];
        int slot = 0;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
#pragma unroll
            for (int b = a; b < 3; ++b) {
                if (!BLOCK && a != b) continue;
                T t[3][3];
                le_coeff(A, AAt_mu, lam_idet, mu_idet, a, b, t);
                const T dsum = (t[0][0] + t[1][1] + t[2][2]) * T(1.0 / 9.0);
                const T mxy = (t[0][1] + t[1][0]) * T(1.0 / 12.0), mxz = (t[0][2] + t[2][0]) * T(1.0 / 12.0),
                        myz = (t[1][2] + t[2][1]) * T(1.0 / 12.0);
#pragma unroll
                for (int v = 0; v < 8; ++v) {
                    const int lex = kVtkToLex[v];
                    const T sx = (lex & 1) ? T(1) : T(-1), sy = (lex & 2) ? T(1) : T(-1), sz = (lex & 4) ? T(1) : T(-1);
                    const T val = dsum + mxy * sx * sy + mxz * sx * sz + myz * sy * sz;
                    if (BLOCK)
                        red_add(&out[(ptrdiff_t)ev[v] * 6 + slot], val);
                    else
                        red_add(&out[(ptrdiff_t)ev[v] * 3 + a], val);
                }
                ++slot;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------

// element operators of the patch kernel (patch.cuh): same arithmetic as the kernels above
template <typename T>
struct Hex8LaplacianLocal {
    const float *fff;  // SoA [6][ne]
    ptrdiff_t ne;
    struct Geo {
        float f[6];
    };
    __device__ __forceinline__ void load(const ptrdiff_t e, Geo &g) const {
#pragma unroll
        for (int k = 0; k < 6; ++k) g.f[k] = ldg_stream(&fff[k * ne + e]);
    }
    __device__ __forceinline__ void compute(const Geo &g, const T (&uin)[8], T (&out)[8]) const {
        T f[6], ul[8], ol[8];
#pragma unroll
        for (int k = 0; k < 6; ++k) f[k] = (T)g.f[k];
#pragma unroll
        for (int v = 0; v < 8; ++v) ul[kVtkToLex[v]] = uin[v];
        hex8_laplacian_local(f, ul, ol);
#pragma unroll
        for (int v = 0; v < 8; ++v) out[v] = ol[kVtkToLex[v]];
    }
};

template <typename T>
struct Hex8LinearElasticityLocal {
    const float *adjugate;     // SoA [9][ne]
    const float *determinant;  // [ne]
    ptrdiff_t ne;
    T mu, lambda;
    struct Geo {
        float a[9], det;
    };
    __device__ __forceinline__ void load(const ptrdiff_t e, Geo &g) const {
#pragma unroll
        for (int k = 0; k < 9; ++k) g.a[k] = ldg_stream(&adjugate[k * ne + e]);
        g.det = ldg_stream(&determinant[e]);
    }
    __device__ __forceinline__ void compute(const Geo &g, const T (&uin)[24], T (&out)[24]) const {
        T A[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) A[k] = (T)g.a[k];
        const T det = (T)g.det;
        T c[3][8], O[3][8];
#pragma unroll
        for (int v = 0; v < 8; ++v)
#pragma unroll
            for (int a = 0; a < 3; ++a) c[a][kVtkToLex[v]] = uin[v * 3 + a];
        Hex8LEConsts<T> K;
        K.init(A, det, T(1), mu, lambda);
        hex8_le_local_apply(K, c, O);
#pragma unroll
        for (int v = 0; v < 8; ++v)
#pragma unroll
            for (int a = 0; a < 3; ++a) out[v * 3 + a] = O[a][kVtkToLex[v]];
    }
};

struct Hex8Laplacian final : OpBase {
    DeviceArray<int32_t> elements;  // [8][ne]
    DeviceArray<float> fff;         // [6][ne]
    ColorPlan colors;
    PatchPlan patch;
    static constexpr int kPatchE = 512;

    template <typename T>
    int apply_t(const T *x, T *y, cudaStream_t s) {
        const int maxgrid = sm_count() * 16;
        if (scatter == B200FEM_SCATTER_PATCH && patch.ready && patch.smem_bytes_direct(1, sizeof(T)) <= 200 * 1024)
            return launch_patch_apply_direct<T, 8, 1, kPatchE, 256, Hex8LaplacianLocal<T>, true, true>(patch, Hex8LaplacianLocal<T>{fff.ptr, ne}, x, y, s);
        if (scatter == B200FEM_SCATTER_COLORED) {
            if (colors.build(ne, nn)) return B200FEM_FAILURE;
            for (int c = 0; c < colors.ncolors(); ++c) {
                const ptrdiff_t b = colors.offsets[c], e = colors.offsets[c + 1];
                if (e == b) continue;
                const int grid = (int)std::min<ptrdiff_t>((e - b + 255) / 256, maxgrid);
                hex8_laplacian_apply_kernel<T, true><<<grid, 256, 0, s>>>(b, e, ne, elements.ptr, fff.ptr, colors.perm.ptr, x, y);
            }
        } else {
            const int grid = (int)std::min<ptrdiff_t>((ne + 255) / 256, maxgrid);
            hex8_laplacian_apply_kernel<T, false><<<grid, 256, 0, s>>>(0, ne, ne, elements.ptr, fff.ptr, nullptr, x, y);
        }
        B2_CHECK_LAUNCH();
        return B200FEM_SUCCESS;
    }
    int apply(const void *x, void *y, bool overwrite, cudaStream_t s) override {
        if (overwrite) B2_CUDA(cudaMemsetAsync(y, 0, (size_t)nn * real_type, s));
        return real_type == B200FEM_F64 ? apply_t<double>((const double *)x, (double *)y, s)
                                        : apply_t<float>((const float *)x, (float *)y, s);
    }
    int diag(void *d, cudaStream_t s) override {
        const int grid = (int)std::min<ptrdiff_t>((ne + 255) / 256, sm_count() * 16);
        if (real_type == B200FEM_F64)
            hex8_laplacian_diag_kernel<double><<<grid, 256, 0, s>>>(ne, elements.ptr, fff.ptr, (double *)d);
        else
            hex8_laplacian_diag_kernel<float><<<grid, 256, 0, s>>>(ne, elements.ptr, fff.ptr, (float *)d);
        B2_CHECK_LAUNCH();
        return B200FEM_SUCCESS;
    }
    int launches(bool overwrite) const override {
        int k = (scatter == B200FEM_SCATTER_COLORED && colors.ready) ? colors.ncolors() : 1;
        return k + (overwrite ? 1 : 0);
    }
};

struct Hex8LinearElasticity final : OpBase {
    DeviceArray<int32_t> elements;  // [8][ne]
    DeviceArray<float> adjugate;    // [9][ne]
    DeviceArray<float> determinant; // [ne]
    double mu = 1, lambda = 1;
    ColorPlan colors;
    PatchPlan patch;
    std::vector<int32_t> host_el;  // SoA copy for the lazily built plans (patch / colours)
    int patch_elements = 256;      // elements per patch: 128, 256 (3 CTAs of 128 threads per SM) or 512

    int ensure_patch() {
        if (patch.ready && patch.E == patch_elements) return B200FEM_SUCCESS;
        if (host_el.empty() && ne > 0) B2_FAIL("hex8 linear elasticity: connectivity no longer on the host");
        std::vector<const int32_t *> rows(8);
        for (int v = 0; v < 8; ++v) rows[v] = host_el.data() + (size_t)v * ne;
        patch.ready = false;
        return patch.build(8, patch_elements, ne, rows.data());
    }
    int set_option(const char *key, int value) override {
        if (!strcmp(key, "patch_elements")) {
            if (value != 128 && value != 256 && value != 512) B2_FAIL("patch_elements must be 128, 256 or 512");
            patch_elements = value;
            return B200FEM_SUCCESS;
        }
        return OpBase::set_option(key, value);
    }
    int get_option(const char *key, int *value) override {
        if (!strcmp(key, "patch_elements")) { *value = patch_elements; return B200FEM_SUCCESS; }
        if (!strcmp(key, "patch_avg_unique_nodes")) { *value = patch.ready ? (int)patch.avg_u : 0; return B200FEM_SUCCESS; }
        return OpBase::get_option(key, value);
    }

    template <typename T>
    int apply_t(const T *x, T *y, cudaStream_t s) {
        const int maxgrid = sm_count() * 16;
        if (scatter == B200FEM_SCATTER_PATCH) {
            if (ensure_patch()) return B200FEM_FAILURE;
            const Hex8LinearElasticityLocal<T> local{adjugate.ptr, determinant.ptr, ne, (T)mu, (T)lambda};
            if (patch.smem_bytes(3, sizeof(T)) <= 200 * 1024) {
                if (patch.E == 128) return launch_patch_apply<T, 8, 3, 128, 128>(patch, local, x, y, s);
                if (patch.E == 256) return launch_patch_apply<T, 8, 3, 256, 128>(patch, local, x, y, s);
                return launch_patch_apply<T, 8, 3, 512, 256>(patch, local, x, y, s);
            }
        }
        if (scatter == B200FEM_SCATTER_COLORED) {
            if (colors.build(ne, nn)) return B200FEM_FAILURE;
            for (int c = 0; c < colors.ncolors(); ++c) {
                const ptrdiff_t b = colors.offsets[c], e = colors.offsets[c + 1];
                if (e == b) continue;
                const int grid = (int)std::min<ptrdiff_t>((e - b + 127) / 128, maxgrid);
                hex8_le_apply_kernel<T, true><<<grid, 128, 0, s>>>(b, e, ne, elements.ptr, adjugate.ptr, determinant.ptr,
                                                                    (T)mu, (T)lambda, colors.perm.ptr, x, y);
            }
        } else {
            const int grid = (int)std::min<ptrdiff_t>((ne + 127) / 128, maxgrid);
            hex8_le_apply_kernel<T, false><<<grid, 128, 0, s>>>(0, ne, ne, elements.ptr, adjugate.ptr, determinant.ptr,
                                                                 (T)mu, (T)lambda, nullptr, x, y);
        }
        B2_CHECK_LAUNCH();
        return B200FEM_SUCCESS;
    }
    int apply(const void *x, void *y, bool overwrite, cudaStream_t s) override {
        if (overwrite) B2_CUDA(cudaMemsetAsync(y, 0, (size_t)nn * 3 * real_type, s));
        return real_type == B200FEM_F64 ? apply_t<double>((const double *)x, (double *)y, s)
                                        : apply_t<float>((const float *)x, (float *)y, s);
    }
    template <bool BLOCK>
    int diag_t(void *d, cudaStream_t s) {
        const int grid = (int)std::min<ptrdiff_t>((ne + 127) / 128, sm_count() * 16);
        if (real_type == B200FEM_F64)
            hex8_le_diag_kernel<double, BLOCK><<<grid, 128, 0, s>>>(ne, elements.ptr, adjugate.ptr, determinant.ptr, mu, lambda, (double *)d);
        else
            hex8_le_diag_kernel<float, BLOCK><<<grid, 128, 0, s>>>(ne, elements.ptr, adjugate.ptr, determinant.ptr, (float)mu, (float)lambda, (float *)d);
        B2_CHECK_LAUNCH();
        return B200FEM_SUCCESS;
    }
    int diag(void *d, cudaStream_t s) override { return diag_t<false>(d, s); }
    int block_diag_sym(void *d, cudaStream_t s) override { return diag_t<true>(d, s); }
    int launches(bool overwrite) const override {
        int k = (scatter == B200FEM_SCATTER_COLORED && colors.ready) ? colors.ncolors() : 1;
        return k + (overwrite ? 1 : 0);
    }
};

static void soa_copy(int nxe, ptrdiff_t ne, const int32_t *const *elements, std::vector<int32_t> &el) {
    el.resize((size_t)nxe * ne);
    for (int v = 0; v < nxe; ++v) std::copy(elements[v], elements[v] + ne, el.begin() + (size_t)v * ne);
}

static void aos_to_soa(int k, ptrdiff_t ne, const float *aos, std::vector<float> &soa) {
    soa.resize((size_t)k * ne);
    for (ptrdiff_t e = 0; e < ne; ++e)
        for (int d = 0; d < k; ++d) soa[(size_t)d * ne + e] = aos[e * k + d];
}

OpBase *make_hex8_laplacian(ptrdiff_t ne, ptrdiff_t nn, const int32_t *const *elements, const float *fff, int real_type) {
    if (!check_indices(8, ne, nn, elements)) {
        set_error("hex8 laplacian: node index out of range");
        return nullptr;
    }
    auto *op = new Hex8Laplacian();
    op->elem_type = B200FEM_HEX8;
    op->real_type = real_type;
    op->ne = ne;
    op->nn = nn;
    op->nxe = 8;
    op->geom_bytes_per_elem = 24;
    // measured at 128^3 (profiles/r01f_extras_patch.jsonl): with 8 elements per node the patch kernel does not beat
    // RED.ADD for hex8 in fp64 (0.354 vs 0.303 ms elasticity, 0.112 vs 0.105 ms Laplacian); it stays selectable
    op->scatter = B200FEM_SCATTER_ATOMIC;
    std::vector<int32_t> el;
    std::vector<float> f;
    soa_copy(8, ne, elements, el);
    aos_to_soa(6, ne, fff, f);
    op->colors.keep(8, ne, elements);
    if (op->patch.build(8, Hex8Laplacian::kPatchE, ne, elements) || op->elements.upload(el.data(), el.size()) ||
        op->fff.upload(f.data(), f.size())) {
        delete op;
        return nullptr;
    }
    return op;
}

namespace {
struct Hex8VectorLaplacian final : OpBase {
    DeviceArray<int32_t> elements;  // [8][ne]
    DeviceArray<float> fff;         // [6][ne]
    template <typename T, bool DIAG>
    int run(const T *x, T *y, cudaStream_t s) {
        if (ne == 0) return B200FEM_SUCCESS;
        const int grid = (int)std::min<ptrdiff_t>((ne + 255) / 256, (ptrdiff_t)sm_count() * 16);
        hex8_vector_laplacian_kernel<T, DIAG><<<grid, 256, 0, s>>>(ne, block_size, elements.ptr, fff.ptr, x, y);
        B2_CHECK_LAUNCH();
        return B200FEM_SUCCESS;
    }
    int apply(const void *x, void *y, bool overwrite, cudaStream_t s) override {
        if (overwrite) B2_CUDA(cudaMemsetAsync(y, 0, (size_t)nn * block_size * real_type, s));
        return real_type == B200FEM_F64 ? run<double, false>((const double *)x, (double *)y, s) : run<float, false>((const float *)x, (float *)y, s);
    }
    int diag(void *d, cudaStream_t s) override {
        return real_type == B200FEM_F64 ? run<double, true>(nullptr, (double *)d, s) : run<float, true>(nullptr, (float *)d, s);
    }
    int launches(bool overwrite) const override { return 1 + (overwrite ? 1 : 0); }
};
}  // namespace

OpBase *make_hex8_vector_laplacian(ptrdiff_t ne, ptrdiff_t nn, const int32_t *const *elements, const float *fff, int vector_size,
                                   int real_type) {
    if (!check_indices(8, ne, nn, elements)) {
        set_error("hex8 vector laplacian: node index out of range");
        return nullptr;
    }
    auto *op = new Hex8VectorLaplacian();
    op->elem_type = B200FEM_HEX8;
    op->real_type = real_type;
    op->block_size = vector_size;
    op->ne = ne;
    op->nn = nn;
    op->nxe = 8;
    op->geom_bytes_per_elem = 24;
    std::vector<int32_t> el;
    std::vector<float> f;
    soa_copy(8, ne, elements, el);
    aos_to_soa(6, ne, fff, f);
    if (op->elements.upload(el.data(), el.size()) || op->fff.upload(f.data(), f.size())) {
        delete op;
        return nullptr;
    }
    return op;
}

OpBase *make_hex8_linear_elasticity(ptrdiff_t ne, ptrdiff_t nn, const int32_t *const *elements, const float *adjugate,
                                    const float *determinant, double mu, double lambda, int real_type) {
    if (!check_indices(8, ne, nn, elements)) {
        set_error("hex8 linear elasticity: node index out of range");
        return nullptr;
    }
    auto *op = new Hex8LinearElasticity();
    op->elem_type = B200FEM_HEX8;
    op->real_type = real_type;
    op->block_size = 3;
    op->ne = ne;
    op->nn = nn;
    op->nxe = 8;
    op->geom_bytes_per_elem = 40;
    // measured at 128^3 (profiles/r02d_c2_sweep.jsonl): the staged patch kernel with 128-element patches beats RED.ADD in
    // both element orders (fp64 0.167 vs 0.243 ms in SFC order, 0.264 vs 0.297 ms in generator order; fp32 0.114 vs 0.181)
    op->scatter = B200FEM_SCATTER_PATCH;
    op->patch_elements = 128;
    op->mu = mu;
    op->lambda = lambda;
    std::vector<int32_t> el;
    std::vector<float> a;
    soa_copy(8, ne, elements, el);
    aos_to_soa(9, ne, adjugate, a);
    op->colors.keep(8, ne, elements);
    op->host_el = el;  // the patch plan is built on first use (scatter = PATCH)
    if (op->elements.upload(el.data(), el.size()) || op->adjugate.upload(a.data(), a.size()) ||
        op->determinant.upload(determinant, (size_t)ne)) {
        delete op;
        return nullptr;
    }
    return op;
}

}  // namespace b200fem
