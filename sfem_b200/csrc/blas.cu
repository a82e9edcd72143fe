// Vector kernels behind sfem::BLAS<T> and the fused CG / Chebyshev updates (include/b200fem.h).
//
// All of them are pure HBM streams. Design for B200:
//   * 16-byte vector loads/stores whenever the pointers allow it, grid = a multiple of the SM count
//     (persistent grid-stride), so every SM keeps >= 8 x 256 x 16 B in flight;
//   * reductions: fp64 accumulation, warp shuffle -> shared -> one partial per block; the LAST block
//     (threadfence + ticket) adds the partials in block order, so a dot is one launch, allocation
//     free, and bit-reproducible for a given n;
//   * results stay in device memory (`*_async`), the blocking variants copy one double through a
//     pinned slot.
#include <math.h>

#include <algorithm>
#include <mutex>
#include <unordered_map>

#include "common.cuh"

namespace b200fem {

static constexpr int kThreads = 256;
static constexpr int kMaxBlocks = 2048;

static inline int grid_for(ptrdiff_t n, int per_thread) {
    ptrdiff_t work = (n + (ptrdiff_t)kThreads * per_thread - 1) / ((ptrdiff_t)kThreads * per_thread);
    ptrdiff_t cap = (ptrdiff_t)sm_count() * 8;
    if (cap > kMaxBlocks) cap = kMaxBlocks;
    if (work < 1) work = 1;
    return (int)(work < cap ? work : cap);
}

template <typename T>
struct Vec;
template <>
struct Vec<double> {
    using type = double2;
    static constexpr int N = 2;
};
template <>
struct Vec<float> {
    using type = float4;
    static constexpr int N = 4;
};

template <typename T>
__device__ __forceinline__ T &lane(typename Vec<T>::type &v, int i) {
    return reinterpret_cast<T *>(&v)[i];
}

static inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// ---------------------------------------------------------------------------------------------
// element-wise kernels: out arrays may alias inputs element-by-element (in-place updates)
// ---------------------------------------------------------------------------------------------

// F: void operator()(ptrdiff_t i) on scalars; vector path calls it on consecutive lanes.
template <typename T, typename F1, typename FV>
__global__ void __launch_bounds__(kThreads) ew_kernel(ptrdiff_t n, bool vec, F1 f1, FV fv) {
    const ptrdiff_t tid = (ptrdiff_t)blockIdx.x * blockDim.x + threadIdx.x;
    const ptrdiff_t stride = (ptrdiff_t)gridDim.x * blockDim.x;
    if (vec) {
        constexpr int N = Vec<T>::N;
        const ptrdiff_t nv = n / N;
        for (ptrdiff_t i = tid; i < nv; i += stride) fv(i);
        for (ptrdiff_t i = nv * N + tid; i < n; i += stride) f1(i);
    } else {
        for (ptrdiff_t i = tid; i < n; i += stride) f1(i);
    }
}

template <typename T, typename F1, typename FV>
static int launch_ew(ptrdiff_t n, bool vec, F1 f1, FV fv, cudaStream_t s) {
    if (n <= 0) return B200FEM_SUCCESS;
    ew_kernel<T><<<grid_for(n, vec ? Vec<T>::N : 1), kThreads, 0, s>>>(n, vec, f1, fv);
    B2_CHECK_LAUNCH();
    return B200FEM_SUCCESS;
}

template <typename T>
static int values_t(ptrdiff_t n, T v, T *x, cudaStream_t s) {
    using V = typename Vec<T>::type;
    return launch_ew<T>(
        n, aligned16(x), [=] __device__(ptrdiff_t i) { x[i] = v; },
        [=] __device__(ptrdiff_t i) {
            V o;
#pragma unroll
            for (int k = 0; k < Vec<T>::N; ++k) lane<T>(o, k) = v;
            reinterpret_cast<V *>(x)[i] = o;
        },
        s);
}

template <typename T>
static int scal_t(ptrdiff_t n, T a, T *x, cudaStream_t s) {
    using V = typename Vec<T>::type;
    return launch_ew<T>(
        n, aligned16(x), [=] __device__(ptrdiff_t i) { x[i] *= a; },
        [=] __device__(ptrdiff_t i) {
            V o = reinterpret_cast<V *>(x)[i];
#pragma unroll
            for (int k = 0; k < Vec<T>::N; ++k) lane<T>(o, k) *= a;
            reinterpret_cast<V *>(x)[i] = o;
        },
        s);
}

template <typename T>
static int reciprocal_t(ptrdiff_t n, T a, T *x, cudaStream_t s) {
    using V = typename Vec<T>::type;
    return launch_ew<T>(
        n, aligned16(x),
        [=] __device__(ptrdiff_t i) {
            T xi = x[i];
            if (xi != T(0)) x[i] = a / xi;
        },
        [=] __device__(ptrdiff_t i) {
            V o = reinterpret_cast<V *>(x)[i];
#pragma unroll
            for (int k = 0; k < Vec<T>::N; ++k) {
                T xi = lane<T>(o, k);
                if (xi != T(0)) lane<T>(o, k) = a / xi;
            }
            reinterpret_cast<V *>(x)[i] = o;
        },
        s);
}

// y = a x + b y      (axpy is b == 1)
template <typename T>
static int axpby_t(ptrdiff_t n, T a, const T *x, T b, T *y, cudaStream_t s) {
    using V = typename Vec<T>::type;
    return launch_ew<T>(
        n, aligned16(x) && aligned16(y), [=] __device__(ptrdiff_t i) { y[i] = a * x[i] + b * y[i]; },
        [=] __device__(ptrdiff_t i) {
            V xv = reinterpret_cast<const V *>(x)[i];
            V yv = reinterpret_cast<V *>(y)[i];
#pragma unroll
            for (int k = 0; k < Vec<T>::N; ++k) lane<T>(yv, k) = a * lane<T>(xv, k) + b * lane<T>(yv, k);
            reinterpret_cast<V *>(y)[i] = yv;
        },
        s);
}

// y += a x  (kept separate from axpby so that y*1 is not rounded differently: it is exact anyway,
// but this saves a multiply and mirrors OpenMP_BLAS::axpy)
template <typename T>
static int axpy_t(ptrdiff_t n, T a, const T *x, T *y, cudaStream_t s) {
    using V = typename Vec<T>::type;
    return launch_ew<T>(
        n, aligned16(x) && aligned16(y), [=] __device__(ptrdiff_t i) { y[i] += a * x[i]; },
        [=] __device__(ptrdiff_t i) {
            V xv = reinterpret_cast<const V *>(x)[i];
            V yv = reinterpret_cast<V *>(y)[i];
#pragma unroll
            for (int k = 0; k < Vec<T>::N; ++k) lane<T>(yv, k) += a * lane<T>(xv, k);
            reinterpret_cast<V *>(y)[i] = yv;
        },
        s);
}

template <typename T>
static int zaxpby_t(ptrdiff_t n, T a, const T *x, T b, const T *y, T *z, cudaStream_t s) {
    using V = typename Vec<T>::type;
    return launch_ew<T>(
        n, aligned16(x) && aligned16(y) && aligned16(z), [=] __device__(ptrdiff_t i) { z[i] = a * x[i] + b * y[i]; },
        [=] __device__(ptrdiff_t i) {
            V xv = reinterpret_cast<const V *>(x)[i];
            V yv = reinterpret_cast<const V *>(y)[i];
            V zv;
#pragma unroll
            for (int k = 0; k < Vec<T>::N; ++k) lane<T>(zv, k) = a * lane<T>(xv, k) + b * lane<T>(yv, k);
            reinterpret_cast<V *>(z)[i] = zv;
        },
        s);
}

template <typename T>
static int xypaz_t(ptrdiff_t n, const T *x, const T *y, T a, T *z, cudaStream_t s) {
    using V = typename Vec<T>::type;
    return launch_ew<T>(
        n, aligned16(x) && aligned16(y) && aligned16(z), [=] __device__(ptrdiff_t i) { z[i] = x[i] * y[i] + a * z[i]; },
        [=] __device__(ptrdiff_t i) {
            V xv = reinterpret_cast<const V *>(x)[i];
            V yv = reinterpret_cast<const V *>(y)[i];
            V zv = reinterpret_cast<V *>(z)[i];
#pragma unroll
            for (int k = 0; k < Vec<T>::N; ++k) lane<T>(zv, k) = lane<T>(xv, k) * lane<T>(yv, k) + a * lane<T>(zv, k);
            reinterpret_cast<V *>(z)[i] = zv;
        },
        s);
}

// p = z + beta p
template <typename T>
static int cg_update_p_t(ptrdiff_t n, T beta, const T *z, T *p, cudaStream_t s) {
    using V = typename Vec<T>::type;
    return launch_ew<T>(
        n, aligned16(z) && aligned16(p), [=] __device__(ptrdiff_t i) { p[i] = z[i] + beta * p[i]; },
        [=] __device__(ptrdiff_t i) {
            V zv = reinterpret_cast<const V *>(z)[i];
            V pv = reinterpret_cast<V *>(p)[i];
#pragma unroll
            for (int k = 0; k < Vec<T>::N; ++k) lane<T>(pv, k) = lane<T>(zv, k) + beta * lane<T>(pv, k);
            reinterpret_cast<V *>(p)[i] = pv;
        },
        s);
}

// p = beta p - rhs + Ax ; x -= alpha p      (Chebyshev3 iteration_update)
template <typename T>
static int cheb3_update_t(ptrdiff_t n, T alpha, T beta, const T *rhs, const T *Ax, T *p, T *x, cudaStream_t s) {
    using V = typename Vec<T>::type;
    return launch_ew<T>(
        n, aligned16(rhs) && aligned16(Ax) && aligned16(p) && aligned16(x),
        [=] __device__(ptrdiff_t i) {
            const T pi = beta * p[i] - rhs[i] + Ax[i];
            p[i] = pi;
            x[i] -= alpha * pi;
        },
        [=] __device__(ptrdiff_t i) {
            V rv = reinterpret_cast<const V *>(rhs)[i];
            V av = reinterpret_cast<const V *>(Ax)[i];
            V pv = reinterpret_cast<V *>(p)[i];
            V xv = reinterpret_cast<V *>(x)[i];
#pragma unroll
            for (int k = 0; k < Vec<T>::N; ++k) {
                const T pi = beta * lane<T>(pv, k) - lane<T>(rv, k) + lane<T>(av, k);
                lane<T>(pv, k) = pi;
                lane<T>(xv, k) -= alpha * pi;
            }
            reinterpret_cast<V *>(p)[i] = pv;
            reinterpret_cast<V *>(x)[i] = xv;
        },
        s);
}

// ---------------------------------------------------------------------------------------------
// reductions
// ---------------------------------------------------------------------------------------------

struct ReduceWorkspace {
    double *partials = nullptr;     // [kMaxBlocks]
    double *result = nullptr;       // device scalar
    unsigned int *ticket = nullptr; // last-block counter (self-resetting)
    double *host_slot = nullptr;    // pinned
};

static std::mutex g_ws_mutex;
static std::unordered_map<uint64_t, ReduceWorkspace> g_ws;

static int get_workspace(cudaStream_t s, ReduceWorkspace **out) {
    int dev = 0;
    B2_CUDA(cudaGetDevice(&dev));
    const uint64_t key = (uint64_t)reinterpret_cast<uintptr_t>(s) * 64u + (uint64_t)dev;
    std::lock_guard<std::mutex> lock(g_ws_mutex);
    auto it = g_ws.find(key);
    if (it == g_ws.end()) {
        ReduceWorkspace w;
        B2_CUDA(cudaMalloc((void **)&w.partials, sizeof(double) * kMaxBlocks));
        B2_CUDA(cudaMalloc((void **)&w.result, sizeof(double)));
        B2_CUDA(cudaMalloc((void **)&w.ticket, sizeof(unsigned int)));
        B2_CUDA(cudaMemset(w.ticket, 0, sizeof(unsigned int)));
        B2_CUDA(cudaMallocHost((void **)&w.host_slot, sizeof(double)));
        it = g_ws.emplace(key, w).first;
    }
    *out = &it->second;
    return B200FEM_SUCCESS;
}

// Block-level sum and deterministic last-block finish. Returns true in thread 0 of the last block
// with *total valid.
__device__ __forceinline__ void block_finish(double acc, double *partials, unsigned int *ticket, double *result) {
    __shared__ double warp_part[kThreads / 32];
    __shared__ bool is_last;
    acc = warp_sum(acc);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) warp_part[w] = acc;
    __syncthreads();
    if (w == 0) {
        double v = (l < kThreads / 32) ? warp_part[l] : 0.0;
        v = warp_sum(v);
        if (l == 0) {
            partials[blockIdx.x] = v;
            __threadfence();
            const unsigned int t = atomicAdd(ticket, 1u);
            is_last = (t == gridDim.x - 1);
        }
    }
    __syncthreads();
    if (is_last) {
        // fixed order: thread t adds partials t, t+256, ...; then the same block tree
        double v = 0.0;
        for (int i = threadIdx.x; i < (int)gridDim.x; i += kThreads) v += __ldcg(&partials[i]);
        v = warp_sum(v);
        if (l == 0) warp_part[w] = v;
        __syncthreads();
        if (w == 0) {
            double t = (l < kThreads / 32) ? warp_part[l] : 0.0;
            t = warp_sum(t);
            if (l == 0) {
                *result = t;
                *ticket = 0u;
            }
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
    dot_kernel(ptrdiff_t n, bool vec, const T *__restrict__ x, const T *__restrict__ y, double *partials,
               unsigned int *ticket, double *result) {
    using V = typename Vec<T>::type;
    const ptrdiff_t tid = (ptrdiff_t)blockIdx.x * blockDim.x + threadIdx.x;
    const ptrdiff_t stride = (ptrdiff_t)gridDim.x * blockDim.x;
    double acc = 0.0;
    if (vec) {
        constexpr int N = Vec<T>::N;
        const ptrdiff_t nv = n / N;
        for (ptrdiff_t i = tid; i < nv; i += stride) {
            V xv = reinterpret_cast<const V *>(x)[i];
            V yv = reinterpret_cast<const V *>(y)[i];
#pragma unroll
            for (int k = 0; k < N; ++k) acc += (double)lane<T>(xv, k) * (double)lane<T>(yv, k);
        }
        for (ptrdiff_t i = nv * N + tid; i < n; i += stride) acc += (double)x[i] * (double)y[i];
    } else {
        for (ptrdiff_t i = tid; i < n; i += stride) acc += (double)x[i] * (double)y[i];
    }
    block_finish(acc, partials, ticket, result);
}

// x += alpha p ; r -= alpha Ap ; result = (r, r)
template <typename T>
__global__ void __launch_bounds__(kThreads)
    cg_xr_kernel(ptrdiff_t n, bool vec, T alpha, const T *__restrict__ p, const T *__restrict__ Ap, T *__restrict__ x,
                 T *__restrict__ r, double *partials, unsigned int *ticket, double *result) {
    using V = typename Vec<T>::type;
    const ptrdiff_t tid = (ptrdiff_t)blockIdx.x * blockDim.x + threadIdx.x;
    const ptrdiff_t stride = (ptrdiff_t)gridDim.x * blockDim.x;
    double acc = 0.0;
    auto one = [&](ptrdiff_t i) {
        x[i] += alpha * p[i];
        const T ri = r[i] - alpha * Ap[i];
        r[i] = ri;
        acc += (double)ri * (double)ri;
    };
    if (vec) {
        constexpr int N = Vec<T>::N;
        const ptrdiff_t nv = n / N;
        for (ptrdiff_t i = tid; i < nv; i += stride) {
            V pv = reinterpret_cast<const V *>(p)[i];
            V av = reinterpret_cast<const V *>(Ap)[i];
            V xv = reinterpret_cast<V *>(x)[i];
            V rv = reinterpret_cast<V *>(r)[i];
#pragma unroll
            for (int k = 0; k < N; ++k) {
                lane<T>(xv, k) += alpha * lane<T>(pv, k);
                const T ri = lane<T>(rv, k) - alpha * lane<T>(av, k);
                lane<T>(rv, k) = ri;
                acc += (double)ri * (double)ri;
            }
            reinterpret_cast<V *>(x)[i] = xv;
            reinterpret_cast<V *>(r)[i] = rv;
        }
        for (ptrdiff_t i = nv * N + tid; i < n; i += stride) one(i);
    } else {
        for (ptrdiff_t i = tid; i < n; i += stride) one(i);
    }
    block_finish(acc, partials, ticket, result);
}

template <typename T>
static int dot_async_t(ptrdiff_t n, const T *x, const T *y, double *result_dev, cudaStream_t s) {
    ReduceWorkspace *w;
    if (get_workspace(s, &w)) return B200FEM_FAILURE;
    if (n <= 0) {
        B2_CUDA(cudaMemsetAsync(result_dev, 0, sizeof(double), s));
        return B200FEM_SUCCESS;
    }
    const bool vec = aligned16(x) && aligned16(y);
    dot_kernel<T><<<grid_for(n, (vec ? Vec<T>::N : 1) * 4), kThreads, 0, s>>>(n, vec, x, y, w->partials, w->ticket,
                                                                               result_dev);
    B2_CHECK_LAUNCH();
    return B200FEM_SUCCESS;
}

template <typename T>
static int dot_sync_t(ptrdiff_t n, const T *x, const T *y, double *result_host, cudaStream_t s) {
    ReduceWorkspace *w;
    if (get_workspace(s, &w)) return B200FEM_FAILURE;
    if (dot_async_t<T>(n, x, y, w->result, s)) return B200FEM_FAILURE;
    B2_CUDA(cudaMemcpyAsync(w->host_slot, w->result, sizeof(double), cudaMemcpyDeviceToHost, s));
    B2_CUDA(cudaStreamSynchronize(s));
    *result_host = *w->host_slot;
    return B200FEM_SUCCESS;
}

template <typename T>
static int cg_xr_async_t(ptrdiff_t n, T alpha, const T *p, const T *Ap, T *x, T *r, double *rtr_dev, cudaStream_t s) {
    ReduceWorkspace *w;
    if (get_workspace(s, &w)) return B200FEM_FAILURE;
    if (n <= 0) {
        B2_CUDA(cudaMemsetAsync(rtr_dev, 0, sizeof(double), s));
        return B200FEM_SUCCESS;
    }
    const bool vec = aligned16(p) && aligned16(Ap) && aligned16(x) && aligned16(r);
    cg_xr_kernel<T><<<grid_for(n, (vec ? Vec<T>::N : 1) * 2), kThreads, 0, s>>>(n, vec, alpha, p, Ap, x, r, w->partials,
                                                                                 w->ticket, rtr_dev);
    B2_CHECK_LAUNCH();
    return B200FEM_SUCCESS;
}

template <typename T>
static int cg_xr_t(ptrdiff_t n, T alpha, const T *p, const T *Ap, T *x, T *r, double *rtr_host, cudaStream_t s) {
    ReduceWorkspace *w;
    if (get_workspace(s, &w)) return B200FEM_FAILURE;
    if (n <= 0) {
        *rtr_host = 0;
        return B200FEM_SUCCESS;
    }
    if (cg_xr_async_t<T>(n, alpha, p, Ap, x, r, w->result, s)) return B200FEM_FAILURE;
    B2_CUDA(cudaMemcpyAsync(w->host_slot, w->result, sizeof(double), cudaMemcpyDeviceToHost, s));
    B2_CUDA(cudaStreamSynchronize(s));
    *rtr_host = *w->host_slot;
    return B200FEM_SUCCESS;
}

// ---------------------------------------------------------------------------------------------
// Dirichlet index-list kernels
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void constraint_copy_kernel(ptrdiff_t n, const int32_t *__restrict__ idx, int bs, int comp,
                                       const T *__restrict__ src, T *__restrict__ dst) {
    for (ptrdiff_t i = (ptrdiff_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (ptrdiff_t)gridDim.x * blockDim.x) {
        const ptrdiff_t d = (ptrdiff_t)idx[i] * bs + comp;
        dst[d] = src[d];
    }
}

template <typename T>
__global__ void constraint_value_kernel(ptrdiff_t n, const int32_t *__restrict__ idx, int bs, int comp, T value,
                                        T *__restrict__ dst) {
    for (ptrdiff_t i = (ptrdiff_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (ptrdiff_t)gridDim.x * blockDim.x) {
        dst[(ptrdiff_t)idx[i] * bs + comp] = value;
    }
}

// ---------------------------------------------------------------------------------------------
// halo pack / unpack (what smesh::Exchange::gather / scatter_add do around their MPI calls)
// ---------------------------------------------------------------------------------------------
template <typename T, bool ADD>
__global__ void __launch_bounds__(256)
    halo_index_kernel(ptrdiff_t n, const int32_t *__restrict__ idx, int bs, const T *__restrict__ src, T *__restrict__ dst) {
    const ptrdiff_t total = n * bs;
    for (ptrdiff_t i = (ptrdiff_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (ptrdiff_t)gridDim.x * blockDim.x) {
        const ptrdiff_t node = i / bs;
        const int c = (int)(i - node * bs);
        if (ADD)
            dst[(ptrdiff_t)idx[node] * bs + c] += src[i];  // unpack-add: indices of one message are distinct
        else
            dst[i] = src[(ptrdiff_t)idx[node] * bs + c];    // pack
    }
}

template <typename T, bool ADD>
static int halo_index_t(ptrdiff_t n, const int32_t *idx, int bs, const T *src, T *dst, cudaStream_t s) {
    if (n <= 0) return B200FEM_SUCCESS;
    const ptrdiff_t total = n * bs;
    const int grid = (int)std::min<ptrdiff_t>((total + 255) / 256, (ptrdiff_t)sm_count() * 8);
    halo_index_kernel<T, ADD><<<grid, 256, 0, s>>>(n, idx, bs, src, dst);
    B2_CHECK_LAUNCH();
    return B200FEM_SUCCESS;
}

}  // namespace b200fem

using namespace b200fem;

#define B2_REAL_DISPATCH(rt, EXPR_D, EXPR_F)                        \
    do {                                                            \
        if ((rt) == B200FEM_F64) return (EXPR_D);                   \
        if ((rt) == B200FEM_F32) return (EXPR_F);                   \
        B2_FAIL("real_type must be B200FEM_F32 or B200FEM_F64");    \
    } while (0)

extern "C" {

int b200fem_blas_zeros(int rt, ptrdiff_t n, void *x, void *stream) {
    if (rt != B200FEM_F32 && rt != B200FEM_F64) B2_FAIL("bad real_type");
    if (n > 0) B2_CUDA(cudaMemsetAsync(x, 0, (size_t)n * rt, as_stream(stream)));
    return B200FEM_SUCCESS;
}

int b200fem_blas_values(int rt, ptrdiff_t n, double v, void *x, void *stream) {
    B2_REAL_DISPATCH(rt, values_t<double>(n, v, (double *)x, as_stream(stream)),
                     values_t<float>(n, (float)v, (float *)x, as_stream(stream)));
}

int b200fem_blas_copy(int rt, ptrdiff_t n, const void *src, void *dest, void *stream) {
    if (rt != B200FEM_F32 && rt != B200FEM_F64) B2_FAIL("bad real_type");
    if (n > 0) B2_CUDA(cudaMemcpyAsync(dest, src, (size_t)n * rt, cudaMemcpyDeviceToDevice, as_stream(stream)));
    return B200FEM_SUCCESS;
}

int b200fem_blas_scal(int rt, ptrdiff_t n, double a, void *x, void *stream) {
    B2_REAL_DISPATCH(rt, scal_t<double>(n, a, (double *)x, as_stream(stream)),
                     scal_t<float>(n, (float)a, (float *)x, as_stream(stream)));
}

int b200fem_blas_axpy(int rt, ptrdiff_t n, double a, const void *x, void *y, void *stream) {
    B2_REAL_DISPATCH(rt, axpy_t<double>(n, a, (const double *)x, (double *)y, as_stream(stream)),
                     axpy_t<float>(n, (float)a, (const float *)x, (float *)y, as_stream(stream)));
}

int b200fem_blas_axpby(int rt, ptrdiff_t n, double a, const void *x, double b, void *y, void *stream) {
    B2_REAL_DISPATCH(rt, axpby_t<double>(n, a, (const double *)x, b, (double *)y, as_stream(stream)),
                     axpby_t<float>(n, (float)a, (const float *)x, (float)b, (float *)y, as_stream(stream)));
}

int b200fem_blas_zaxpby(int rt, ptrdiff_t n, double a, const void *x, double b, const void *y, void *z, void *stream) {
    B2_REAL_DISPATCH(
        rt, zaxpby_t<double>(n, a, (const double *)x, b, (const double *)y, (double *)z, as_stream(stream)),
        zaxpby_t<float>(n, (float)a, (const float *)x, (float)b, (const float *)y, (float *)z, as_stream(stream)));
}

int b200fem_blas_xypaz(int rt, ptrdiff_t n, const void *x, const void *y, double a, void *z, void *stream) {
    B2_REAL_DISPATCH(rt, xypaz_t<double>(n, (const double *)x, (const double *)y, a, (double *)z, as_stream(stream)),
                     xypaz_t<float>(n, (const float *)x, (const float *)y, (float)a, (float *)z, as_stream(stream)));
}

int b200fem_blas_reciprocal(int rt, ptrdiff_t n, double a, void *x, void *stream) {
    B2_REAL_DISPATCH(rt, reciprocal_t<double>(n, a, (double *)x, as_stream(stream)),
                     reciprocal_t<float>(n, (float)a, (float *)x, as_stream(stream)));
}

int b200fem_blas_dot(int rt, ptrdiff_t n, const void *x, const void *y, double *result_host, void *stream) {
    if (!result_host) B2_FAIL("dot: NULL result");
    B2_REAL_DISPATCH(rt, dot_sync_t<double>(n, (const double *)x, (const double *)y, result_host, as_stream(stream)),
                     dot_sync_t<float>(n, (const float *)x, (const float *)y, result_host, as_stream(stream)));
}

int b200fem_blas_norm2(int rt, ptrdiff_t n, const void *x, double *result_host, void *stream) {
    if (b200fem_blas_dot(rt, n, x, x, result_host, stream)) return B200FEM_FAILURE;
    *result_host = sqrt(*result_host);
    return B200FEM_SUCCESS;
}

int b200fem_blas_dot_async(int rt, ptrdiff_t n, const void *x, const void *y, double *result_dev, void *stream) {
    if (!result_dev) B2_FAIL("dot_async: NULL result");
    B2_REAL_DISPATCH(rt, dot_async_t<double>(n, (const double *)x, (const double *)y, result_dev, as_stream(stream)),
                     dot_async_t<float>(n, (const float *)x, (const float *)y, result_dev, as_stream(stream)));
}

int b200fem_cg_update_x_r_rtr(int rt, ptrdiff_t n, double alpha, const void *p, const void *Ap, void *x, void *r,
                              double *rtr_host, void *stream) {
    if (!rtr_host) B2_FAIL("cg_update_x_r_rtr: NULL result");
    B2_REAL_DISPATCH(
        rt, cg_xr_t<double>(n, alpha, (const double *)p, (const double *)Ap, (double *)x, (double *)r, rtr_host, as_stream(stream)),
        cg_xr_t<float>(n, (float)alpha, (const float *)p, (const float *)Ap, (float *)x, (float *)r, rtr_host, as_stream(stream)));
}

int b200fem_cg_update_x_r_rtr_async(int rt, ptrdiff_t n, double alpha, const void *p, const void *Ap, void *x, void *r,
                                    double *rtr_dev, void *stream) {
    if (!rtr_dev) B2_FAIL("cg_update_x_r_rtr_async: NULL result");
    B2_REAL_DISPATCH(
        rt, cg_xr_async_t<double>(n, alpha, (const double *)p, (const double *)Ap, (double *)x, (double *)r, rtr_dev, as_stream(stream)),
        cg_xr_async_t<float>(n, (float)alpha, (const float *)p, (const float *)Ap, (float *)x, (float *)r, rtr_dev, as_stream(stream)));
}

int b200fem_cg_update_p(int rt, ptrdiff_t n, double beta, const void *z, void *p, void *stream) {
    B2_REAL_DISPATCH(rt, cg_update_p_t<double>(n, beta, (const double *)z, (double *)p, as_stream(stream)),
                     cg_update_p_t<float>(n, (float)beta, (const float *)z, (float *)p, as_stream(stream)));
}

int b200fem_cheb3_update(int rt, ptrdiff_t n, double alpha, double beta, const void *rhs, const void *Ax, void *p,
                         void *x, void *stream) {
    B2_REAL_DISPATCH(rt,
                     cheb3_update_t<double>(n, alpha, beta, (const double *)rhs, (const double *)Ax, (double *)p,
                                            (double *)x, as_stream(stream)),
                     cheb3_update_t<float>(n, (float)alpha, (float)beta, (const float *)rhs, (const float *)Ax,
                                           (float *)p, (float *)x, as_stream(stream)));
}

int b200fem_constraint_nodes_copy_vec(int rt, ptrdiff_t n, const int32_t *idx, int bs, int comp, const void *src,
                                      void *dst, void *stream) {
    if (n <= 0) return B200FEM_SUCCESS;
    const int grid = (int)std::min<ptrdiff_t>((n + 255) / 256, (ptrdiff_t)sm_count() * 8);
    if (rt == B200FEM_F64)
        constraint_copy_kernel<double><<<grid, 256, 0, as_stream(stream)>>>(n, idx, bs, comp, (const double *)src, (double *)dst);
    else if (rt == B200FEM_F32)
        constraint_copy_kernel<float><<<grid, 256, 0, as_stream(stream)>>>(n, idx, bs, comp, (const float *)src, (float *)dst);
    else
        B2_FAIL("bad real_type");
    B2_CHECK_LAUNCH();
    return B200FEM_SUCCESS;
}

int b200fem_constraint_nodes_to_value_vec(int rt, ptrdiff_t n, const int32_t *idx, int bs, int comp, double value,
                                          void *values, void *stream) {
    if (n <= 0) return B200FEM_SUCCESS;
    const int grid = (int)std::min<ptrdiff_t>((n + 255) / 256, (ptrdiff_t)sm_count() * 8);
    if (rt == B200FEM_F64)
        constraint_value_kernel<double><<<grid, 256, 0, as_stream(stream)>>>(n, idx, bs, comp, value, (double *)values);
    else if (rt == B200FEM_F32)
        constraint_value_kernel<float><<<grid, 256, 0, as_stream(stream)>>>(n, idx, bs, comp, (float)value, (float *)values);
    else
        B2_FAIL("bad real_type");
    B2_CHECK_LAUNCH();
    return B200FEM_SUCCESS;
}

int b200fem_halo_pack(int rt, ptrdiff_t n, const int32_t *idx, int bs, const void *src, void *dst, void *stream) {
    if (bs < 1) B2_FAIL("halo_pack: block_size must be >= 1");
    B2_REAL_DISPATCH(rt, (halo_index_t<double, false>(n, idx, bs, (const double *)src, (double *)dst, as_stream(stream))),
                     (halo_index_t<float, false>(n, idx, bs, (const float *)src, (float *)dst, as_stream(stream))));
}

int b200fem_halo_unpack_add(int rt, ptrdiff_t n, const int32_t *idx, int bs, const void *src, void *dst, void *stream) {
    if (bs < 1) B2_FAIL("halo_unpack_add: block_size must be >= 1");
    B2_REAL_DISPATCH(rt, (halo_index_t<double, true>(n, idx, bs, (const double *)src, (double *)dst, as_stream(stream))),
                     (halo_index_t<float, true>(n, idx, bs, (const float *)src, (float *)dst, as_stream(stream))));
}

}  // extern "C"
