// Semi-structured neighbours of the sshex8 Laplacian:
//
//   * VectorLaplacian on sshex8: affine_sshex8_vector_laplacian_apply_fff (operators/sshex8/sshex8_vector_laplacian.cpp:16-97;
//     VectorLaplacian::apply -> vector_laplacian_apply_opt, frontend/ops/sfem_VectorLaplacian.cpp:191-206,
//     operators/vector_laplacian.cpp:41-66) -- the scalar stencil operator on every component of an AoS field.
//     Built on the scalar sshex8 plan (the warp-specialised stencil kernels work on unit-stride vectors): component b is
//     de-interleaved into a plan-owned work vector, the scalar kernel writes a second one, and one pass adds it back
//     into the AoS result. Two extra vector passes per component next to the scalar apply; nothing is allocated in apply.
//   * LumpedMass on sshex8: affine_sshex8_mass_lumped (operators/sshex8/sshex8_mass.cpp:11-115; dispatch
//     operators/mass.cpp:74-77): the 27-point lumped mass of the MACRO hex scaled by h^3 on every micro element.
#include <algorithm>
#include <vector>

#include "coloring.cuh"
#include "common.cuh"
#include "hex8_local.cuh"
#include "hex8_mass_local.cuh"

namespace b200fem {

template <typename T>
__global__ void __launch_bounds__(256) component_gather_kernel(const ptrdiff_t n, const int vs, const int b, const T *__restrict__ aos, T *__restrict__ out) {
    for (ptrdiff_t i = (ptrdiff_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (ptrdiff_t)gridDim.x * blockDim.x) out[i] = aos[i * vs + b];
}
template <typename T, bool OVERWRITE>
__global__ void __launch_bounds__(256) component_scatter_kernel(const ptrdiff_t n, const int vs, const int b, const T *__restrict__ in, T *__restrict__ aos) {
    for (ptrdiff_t i = (ptrdiff_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (ptrdiff_t)gridDim.x * blockDim.x) {
        if (OVERWRITE)
            aos[i * vs + b] = in[i];
        else
            aos[i * vs + b] += in[i];
    }
}
// every component takes the same diagonal
template <typename T>
__global__ void __launch_bounds__(256) component_broadcast_add_kernel(const ptrdiff_t n, const int vs, const T *__restrict__ in, T *__restrict__ aos) {
    for (ptrdiff_t i = (ptrdiff_t)blockIdx.x * blockDim.x + threadIdx.x; i < n * vs; i += (ptrdiff_t)gridDim.x * blockDim.x) aos[i] += in[i / vs];
}

struct Sshex8VectorLaplacian final : OpBase {
    OpBase *scalar = nullptr;
    DeviceArray<unsigned char> work;  // 2 x nn scalars
    ~Sshex8VectorLaplacian() override { delete scalar; }

    template <typename T>
    int apply_t(const T *x, T *y, bool overwrite, cudaStream_t s) {
        if (nn == 0) return B200FEM_SUCCESS;
        T *xs = reinterpret_cast<T *>(work.ptr), *ys = xs + nn;
        const int grid = (int)std::min<ptrdiff_t>((nn + 255) / 256, (ptrdiff_t)sm_count() * 16);
        for (int b = 0; b < block_size; ++b) {
            component_gather_kernel<T><<<grid, 256, 0, s>>>(nn, block_size, b, x, xs);
            B2_CHECK_LAUNCH();
            if (scalar->apply(xs, ys, true, s)) return B200FEM_FAILURE;
            if (overwrite)
                component_scatter_kernel<T, true><<<grid, 256, 0, s>>>(nn, block_size, b, ys, y);
            else
                component_scatter_kernel<T, false><<<grid, 256, 0, s>>>(nn, block_size, b, ys, y);
            B2_CHECK_LAUNCH();
        }
        return B200FEM_SUCCESS;
    }
    int apply(const void *x, void *y, bool overwrite, cudaStream_t s) override {
        return real_type == B200FEM_F64 ? apply_t<double>((const double *)x, (double *)y, overwrite, s)
                                        : apply_t<float>((const float *)x, (float *)y, overwrite, s);
    }
    template <typename T>
    int diag_t(T *d, cudaStream_t s) {
        if (nn == 0) return B200FEM_SUCCESS;
        T *ds = reinterpret_cast<T *>(work.ptr);
        B2_CUDA(cudaMemsetAsync(ds, 0, (size_t)nn * sizeof(T), s));
        if (scalar->diag(ds, s)) return B200FEM_FAILURE;
        const int grid = (int)std::min<ptrdiff_t>((nn * block_size + 255) / 256, (ptrdiff_t)sm_count() * 16);
        component_broadcast_add_kernel<T><<<grid, 256, 0, s>>>(nn, block_size, ds, d);
        B2_CHECK_LAUNCH();
        return B200FEM_SUCCESS;
    }
    int diag(void *d, cudaStream_t s) override { return real_type == B200FEM_F64 ? diag_t((double *)d, s) : diag_t((float *)d, s); }
    int launches(bool overwrite) const override {
        (void)overwrite;  // the scalar kernel always writes the work vector
        return block_size * (2 + scalar->launches(true));
    }
    double streamed_bytes(bool overwrite) const override {
        (void)overwrite;
        return block_size * (scalar->streamed_bytes(true) + 4.0 * double(nn) * real_type);
    }
    int set_option(const char *key, int value) override { return scalar->set_option(key, value); }
    int get_option(const char *key, int *value) override { return scalar->get_option(key, value); }
};

OpBase *make_sshex8_vector_laplacian(int level, ptrdiff_t ne, ptrdiff_t nn, const int32_t *const *elements, const float *fff, int vector_size,
                                     int real_type) {
    OpBase *scalar = make_sshex8_laplacian(level, ne, nn, elements, fff, real_type);
    if (!scalar) return nullptr;
    auto *op = new Sshex8VectorLaplacian();
    op->scalar = scalar;
    op->elem_type = B200FEM_SSHEX8;
    op->level = level;
    op->real_type = real_type;
    op->block_size = vector_size;
    op->ne = ne;
    op->nn = nn;
    op->nxe = scalar->nxe;
    op->geom_bytes_per_elem = scalar->geom_bytes_per_elem;
    if (op->work.alloc((size_t)2 * nn * real_type)) {
        delete op;
        return nullptr;
    }
    return op;
}

// ---------------------------------------------------------------------------------------------- lumped mass
// one thread per macro element: lumped mass of the macro hex (27-point Gauss, isoparametric) times h^3 = the lumped mass of every
// micro element; lattice node (x, y, z) collects corner (bx, by, bz) of the micro element (x - bx, y - by, z - bz) where that exists
template <typename T>
__global__ void __launch_bounds__(128)
    sshex8_lumped_mass_kernel(const int L, const ptrdiff_t ne, const ptrdiff_t nn, const int32_t *__restrict__ elements,
                              const float *__restrict__ points, T *__restrict__ diag) {
    const int N = L + 1;
    for (ptrdiff_t e = (ptrdiff_t)blockIdx.x * blockDim.x + threadIdx.x; e < ne; e += (ptrdiff_t)gridDim.x * blockDim.x) {
        T X[3][8], U[8], O[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int bx = i & 1, by = (i >> 1) & 1, bz = (i >> 2) & 1;
            const ptrdiff_t node = elements[(size_t)((bz * L * N + by * L) * N + bx * L) * ne + e];
#pragma unroll
            for (int d = 0; d < 3; ++d) X[d][i] = (T)points[(size_t)d * nn + node];
            U[i] = T(1);
        }
        hex8_mass_local(X, U, O);
        const T h = T(1) / T(L), h3 = h * h * h;
#pragma unroll
        for (int i = 0; i < 8; ++i) O[i] *= h3;
        for (int z = 0; z < N; ++z)
            for (int y = 0; y < N; ++y)
                for (int x = 0; x < N; ++x) {
                    T acc = T(0);
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int mx = x - (i & 1), my = y - ((i >> 1) & 1), mz = z - ((i >> 2) & 1);
                        if (mx >= 0 && mx < L && my >= 0 && my < L && mz >= 0 && mz < L) acc += O[i];
                    }
                    red_add(&diag[elements[(size_t)((z * N + y) * N + x) * ne + e]], acc);
                }
    }
}

struct Sshex8LumpedMass final : OpBase {
    DeviceArray<int32_t> elements;  // [(L+1)^3][ne]
    DeviceArray<float> points;      // [3][nn]
    int apply(const void *, void *, bool, cudaStream_t) override {
        set_error("sshex8 mass: lumped only (the reference's apply_mass has no semi-structured case, operators/mass.cpp:158-207)");
        return B200FEM_FAILURE;
    }
    int diag(void *d, cudaStream_t s) override {
        if (ne == 0) return B200FEM_SUCCESS;
        const int grid = (int)std::min<ptrdiff_t>((ne + 127) / 128, (ptrdiff_t)sm_count() * 16);
        if (real_type == B200FEM_F64)
            sshex8_lumped_mass_kernel<double><<<grid, 128, 0, s>>>(level, ne, nn, elements.ptr, points.ptr, (double *)d);
        else
            sshex8_lumped_mass_kernel<float><<<grid, 128, 0, s>>>(level, ne, nn, elements.ptr, points.ptr, (float *)d);
        B2_CHECK_LAUNCH();
        return B200FEM_SUCCESS;
    }
};

OpBase *make_sshex8_lumped_mass(int level, ptrdiff_t ne, ptrdiff_t nn, const int32_t *const *elements, const float *const *points,
                                int real_type) {
    const int nxe = (level + 1) * (level + 1) * (level + 1);
    if (!check_indices(nxe, ne, nn, elements)) {
        set_error("sshex8 lumped mass: node index out of range");
        return nullptr;
    }
    auto *op = new Sshex8LumpedMass();
    op->elem_type = B200FEM_SSHEX8;
    op->level = level;
    op->real_type = real_type;
    op->ne = ne;
    op->nn = nn;
    op->nxe = nxe;
    op->geom_bytes_per_elem = 96;
    std::vector<int32_t> el((size_t)nxe * ne);
    for (int v = 0; v < nxe; ++v) std::copy(elements[v], elements[v] + ne, el.begin() + (size_t)v * ne);
    bool bad = op->elements.upload(el.data(), el.size()) != 0 || op->points.alloc((size_t)3 * nn) != 0;
    for (int d = 0; d < 3 && !bad; ++d)
        bad = cudaMemcpy(op->points.ptr + (size_t)d * nn, points[d], (size_t)nn * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess;
    if (bad) {
        delete op;
        return nullptr;
    }
    return op;
}

}  // namespace b200fem
