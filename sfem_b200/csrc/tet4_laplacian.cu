// tet4 Laplacian, matrix-free:  y[ev] += K_e(fff) u[ev]
//
// Replaces tet4_laplacian_apply_opt / tet4_laplacian_diag_opt
// (operators/tet4/tet4_laplacian.cpp:237-322 of the reference; local operator
// tet4_laplacian_apply_fff, operators/tet4/tet4_laplacian_inline_cpu.hpp:54-73).
//
// Local operator restated: with FFF = [a b c; b d e; c e f] (= J^-1 J^-T det/6, 6 unique fp32
// entries widened to T) and reference gradients g0 = (-1,-1,-1), g1 = e_x, g2 = e_y, g3 = e_z,
// the element vector is  out_i = g_i . (FFF * grad_u),  grad_u = sum_j g_j u_j
// = (u1-u0, u2-u0, u3-u0).  So: w = FFF * (u1-u0, u2-u0, u3-u0); out = (-(w0+w1+w2), w0, w1, w2).
// 15 flops instead of the ~45 of the expanded form; identical in exact arithmetic.
//
// Data layout in HBM (plan): connectivity SoA int32 [4][ne] (coalesced over elements), FFF
// re-laid-out SoA fp32 [6][ne]; both streamed once per apply with evict-first loads. u/y gathers and
// the 4 scatter-adds per element hit L2 (vectors are small next to 40 B/element of streams).
// One thread per element; scatter = RED.ADD (B200FEM_SCATTER_ATOMIC) or element colouring.
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "coloring.cuh"
#include "common.cuh"
#include "patch.cuh"

namespace b200fem {

template <typename T, bool COLORED>
__global__ void __launch_bounds__(256)
    tet4_laplacian_apply_kernel(const ptrdiff_t begin, const ptrdiff_t end, const ptrdiff_t ne,
                                const int32_t *__restrict__ elements, const float *__restrict__ fff,
                                const int32_t *__restrict__ perm, const T *__restrict__ u, T *__restrict__ y) {
    for (ptrdiff_t i = begin + (ptrdiff_t)blockIdx.x * blockDim.x + threadIdx.x; i < end;
         i += (ptrdiff_t)gridDim.x * blockDim.x) {
        const ptrdiff_t e = COLORED ? (ptrdiff_t)perm[i] : i;
        int32_t ev[4];
#pragma unroll
        for (int v = 0; v < 4; ++v) ev[v] = COLORED ? __ldg(&elements[v * ne + e]) : ldg_stream(&elements[v * ne + e]);
        T f[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) f[k] = (T)(COLORED ? __ldg(&fff[k * ne + e]) : ldg_stream(&fff[k * ne + e]));

        const T u0 = u[ev[0]];
        const T d1 = u[ev[1]] - u0, d2 = u[ev[2]] - u0, d3 = u[ev[3]] - u0;
        const T w0 = f[0] * d1 + f[1] * d2 + f[2] * d3;
        const T w1 = f[1] * d1 + f[3] * d2 + f[4] * d3;
        const T w2 = f[2] * d1 + f[4] * d2 + f[5] * d3;
        const T o0 = -(w0 + w1 + w2);
        if (COLORED) {
            y[ev[0]] += o0;
            y[ev[1]] += w0;
            y[ev[2]] += w1;
            y[ev[3]] += w2;
        } else {
            red_add(&y[ev[0]], o0);
            red_add(&y[ev[1]], w0);
            red_add(&y[ev[2]], w1);
            red_add(&y[ev[3]], w2);
        }
    }
}

// element operator of the patch kernel (patch.cuh): same arithmetic as above
template <typename T>
struct Tet4LaplacianLocal {
    const float *fff;  // SoA [6][ne]
    ptrdiff_t ne;
    struct Geo {
        float f[6];
    };
    __device__ __forceinline__ void load(const ptrdiff_t e, Geo &g) const {
#pragma unroll
        for (int k = 0; k < 6; ++k) g.f[k] = ldg_stream(&fff[k * ne + e]);
    }
    __device__ __forceinline__ void compute(const Geo &g, const T (&uin)[4], T (&out)[4]) const {
        T f[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) f[k] = (T)g.f[k];
        const T d1 = uin[1] - uin[0], d2 = uin[2] - uin[0], d3 = uin[3] - uin[0];
        const T w0 = f[0] * d1 + f[1] * d2 + f[2] * d3;
        const T w1 = f[1] * d1 + f[3] * d2 + f[4] * d3;
        const T w2 = f[2] * d1 + f[4] * d2 + f[5] * d3;
        out[0] = -(w0 + w1 + w2);
        out[1] = w0;
        out[2] = w1;
        out[3] = w2;
    }
};

template <typename T>
__global__ void __launch_bounds__(256)
    tet4_laplacian_diag_kernel(const ptrdiff_t ne, const int32_t *__restrict__ elements,
                               const float *__restrict__ fff, T *__restrict__ diag) {
    for (ptrdiff_t e = (ptrdiff_t)blockIdx.x * blockDim.x + threadIdx.x; e < ne; e += (ptrdiff_t)gridDim.x * blockDim.x) {
        T f[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) f[k] = (T)fff[k * ne + e];
        // diagonal of the 4x4: g_i . FFF g_i
        red_add(&diag[elements[0 * ne + e]], f[0] + 2 * f[1] + 2 * f[2] + f[3] + 2 * f[4] + f[5]);
        red_add(&diag[elements[1 * ne + e]], f[0]);
        red_add(&diag[elements[2 * ne + e]], f[3]);
        red_add(&diag[elements[3 * ne + e]], f[5]);
    }
}

struct Tet4Laplacian final : OpBase {
    DeviceArray<int32_t> elements;  // [4][ne]
    DeviceArray<float> fff;         // [6][ne]
    ColorPlan colors;
    PatchPlan patch;
    int patch_pipelined = 3;    // 0 = direct, 2 = direct with the incidence lists staged by cp.async (1.45 vs 1.53 ms at 260^3),
                                // 3 = 2 + the element loads issued before the node gather (default: 1.37 ms), 1 = fully
                                // software-pipelined kernel (2 CTAs per SM: slower for this light operator)
    int patch_smem_atomic = 0;  // 1 = shared-memory atomics instead of the incidence lists (order not fixed)
    static constexpr int kPatchE = 1024;

    template <typename T>
    int apply_t(const T *x, T *y, cudaStream_t s) {
        const int maxgrid = sm_count() * 16;
        if (scatter == B200FEM_SCATTER_PATCH && patch.ready && patch.smem_bytes_direct(1, sizeof(T)) <= 200 * 1024) {
            if (patch.E == 256)  // small meshes: 256-tet patches, 8 CTAs per SM (see make_tet4_laplacian)
                return launch_patch_apply_direct<T, 4, 1, 256, 256, Tet4LaplacianLocal<T>, true, true>(patch, Tet4LaplacianLocal<T>{fff.ptr, ne}, x, y, s);
            if (patch_smem_atomic)
                return launch_patch_apply_smem_atomic<T, 4, 1, kPatchE, 256>(patch, Tet4LaplacianLocal<T>{fff.ptr, ne}, x, y, s);
            if (patch_pipelined == 1) return launch_patch_apply<T, 4, 1, kPatchE, 256>(patch, Tet4LaplacianLocal<T>{fff.ptr, ne}, x, y, s);
            if (patch_pipelined == 2)
                return launch_patch_apply_direct<T, 4, 1, kPatchE, 256, Tet4LaplacianLocal<T>, true>(patch, Tet4LaplacianLocal<T>{fff.ptr, ne}, x, y, s);
            if (patch_pipelined == 3)  // + the element loads issued before the node gather
                return launch_patch_apply_direct<T, 4, 1, kPatchE, 256, Tet4LaplacianLocal<T>, true, true>(patch, Tet4LaplacianLocal<T>{fff.ptr, ne}, x, y, s);
            return launch_patch_apply_direct<T, 4, 1, kPatchE, 256>(patch, Tet4LaplacianLocal<T>{fff.ptr, ne}, x, y, s);
        }
        if (scatter == B200FEM_SCATTER_COLORED) {
            if (colors.build(ne, nn)) return B200FEM_FAILURE;
            for (int c = 0; c < colors.ncolors(); ++c) {
                const ptrdiff_t b = colors.offsets[c], e = colors.offsets[c + 1];
                if (e == b) continue;
                const int grid = (int)std::min<ptrdiff_t>((e - b + 255) / 256, maxgrid);
                tet4_laplacian_apply_kernel<T, true><<<grid, 256, 0, s>>>(b, e, ne, elements.ptr, fff.ptr, colors.perm.ptr, x, y);
            }
        } else {
            const int grid = (int)std::min<ptrdiff_t>((ne + 255) / 256, maxgrid);
            tet4_laplacian_apply_kernel<T, false><<<grid, 256, 0, s>>>(0, ne, ne, elements.ptr, fff.ptr, nullptr, x, y);
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
            tet4_laplacian_diag_kernel<double><<<grid, 256, 0, s>>>(ne, elements.ptr, fff.ptr, (double *)d);
        else
            tet4_laplacian_diag_kernel<float><<<grid, 256, 0, s>>>(ne, elements.ptr, fff.ptr, (float *)d);
        B2_CHECK_LAUNCH();
        return B200FEM_SUCCESS;
    }

    int launches(bool overwrite) const override {
        int k = (scatter == B200FEM_SCATTER_COLORED && colors.ready) ? colors.ncolors() : 1;
        return k + (overwrite ? 1 : 0);
    }
    double streamed_bytes(bool overwrite) const override {
        const double idx = (scatter == B200FEM_SCATTER_PATCH && patch.ready) ? patch.index_bytes() : 16.0 * ne;
        return idx + 24.0 * ne + double(nn) * real_type * (overwrite ? 2.0 : 3.0);
    }
    int set_option(const char *key, int value) override {
        if (!strcmp(key, "patch_smem_atomic")) { patch_smem_atomic = value; return B200FEM_SUCCESS; }
        if (!strcmp(key, "patch_pipelined")) { patch_pipelined = value; return B200FEM_SUCCESS; }
        return OpBase::set_option(key, value);
    }
    int get_option(const char *key, int *value) override {
        if (!strcmp(key, "patch_avg_unique_nodes")) { *value = (int)patch.avg_u; return B200FEM_SUCCESS; }
        return OpBase::get_option(key, value);
    }
};

OpBase *make_tet4_laplacian(ptrdiff_t ne, ptrdiff_t nn, const int32_t *const *elements, const float *fff, int real_type) {
    auto *op = new Tet4Laplacian();
    op->elem_type = B200FEM_TET4;
    op->real_type = real_type;
    op->ne = ne;
    op->nn = nn;
    op->nxe = 4;
    op->geom_bytes_per_elem = 24;
    op->scatter = B200FEM_SCATTER_PATCH;
    if (!check_indices(4, ne, nn, elements)) {
        set_error("tet4 laplacian: node index out of range");
        delete op;
        return nullptr;
    }
    std::vector<int32_t> el((size_t)4 * ne);
    for (int v = 0; v < 4; ++v) std::copy(elements[v], elements[v] + ne, el.begin() + (size_t)v * ne);
    op->colors.keep(4, ne, elements);
    std::vector<float> f((size_t)6 * ne);
    for (ptrdiff_t e = 0; e < ne; ++e)
        for (int k = 0; k < 6; ++k) f[(size_t)k * ne + e] = fff[e * 6 + k];
    // patch size: 1024 tets amortise a patch's unique nodes best (0.45 nodes per tet) and win on large meshes; a mesh of a few
    // million tets is only a handful of such patches per SM and the apply is bound by the latency chain of ONE patch (C1, 1e6
    // tets: 22 - 26 us for 5 us of bytes), so small meshes take 256-tet patches, 8 CTAs per SM
    int patch_e = ne <= (ptrdiff_t)4000000 ? 256 : Tet4Laplacian::kPatchE;
    if (const char *env = getenv("B200FEM_TET4_PATCH_E")) patch_e = atoi(env) == 256 ? 256 : Tet4Laplacian::kPatchE;
    if (op->patch.build(4, patch_e, ne, elements) || op->elements.upload(el.data(), el.size()) ||
        op->fff.upload(f.data(), f.size())) {
        delete op;
        return nullptr;
    }
    return op;
}

}  // namespace b200fem
