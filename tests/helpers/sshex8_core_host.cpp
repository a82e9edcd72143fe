// CPU harness for sfem_b200/csrc/sshex8_laplacian_core.cuh: runs the exact per-thread code of the
// CUDA kernel (one "thread" per lattice row, barrier-separated read/write phases, in-place output)
// sequentially on the host so the formulation can be checked against the oracle without a GPU.
#include <stddef.h>
#include <stdint.h>

#include <vector>

#include "sshex8_laplacian_core.cuh"

using namespace b200fem;

template <typename T, int L>
static void run(ptrdiff_t ne, const int32_t *const *elements, const float *fff, const T *u, T *out) {
    constexpr int N = L + 1;
    std::vector<T> tile(N * N * N);
    std::vector<Sshex8LaplacianRow<T, L>> rows(N);
    for (ptrdiff_t e = 0; e < ne; ++e) {
        for (int n = 0; n < N * N * N; ++n) tile[n] = u[elements[n][e]];
        T f[6];
        for (int k = 0; k < 6; ++k) f[k] = (T)fff[e * 6 + k];
        for (int y = 0; y < N; ++y) rows[y].init(f, y);
        for (int zp = 0; zp <= L; ++zp) {
            for (int y = 0; y < N; ++y) rows[y].step(zp, &tile[zp * N * N], N);
            // barrier
            for (int y = 0; y < N; ++y) rows[y].write_and_rotate(zp, &tile[(zp > 0 ? zp - 1 : 0) * N * N], N);
        }
        for (int y = 0; y < N; ++y) rows[y].write_last(&tile[L * N * N], N);
        for (int n = 0; n < N * N * N; ++n) out[elements[n][e]] += tile[n];
    }
}

template <typename T>
static int dispatch(int level, ptrdiff_t ne, const int32_t *const *elements, const float *fff, const T *u, T *out) {
    switch (level) {
        case 2: run<T, 2>(ne, elements, fff, u, out); return 0;
        case 3: run<T, 3>(ne, elements, fff, u, out); return 0;
        case 4: run<T, 4>(ne, elements, fff, u, out); return 0;
        case 8: run<T, 8>(ne, elements, fff, u, out); return 0;
        default: return 1;
    }
}

extern "C" int sshex8_core_apply_f64(int level, ptrdiff_t ne, const int32_t *const *elements, const float *fff,
                                     const double *u, double *out) {
    return dispatch<double>(level, ne, elements, fff, u, out);
}
extern "C" int sshex8_core_apply_f32(int level, ptrdiff_t ne, const int32_t *const *elements, const float *fff,
                                     const float *u, float *out) {
    return dispatch<float>(level, ne, elements, fff, u, out);
}
