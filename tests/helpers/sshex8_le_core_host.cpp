// CPU harness for sfem_b200/csrc/hex8_le_core.cuh: the exact per-micro-element arithmetic of the sshex8
// linear-elasticity CUDA kernel, run sequentially on the host over the lattice of every macro element, so
// the formulation can be checked against the oracle without a GPU.
#include <stddef.h>
#include <stdint.h>

#include <vector>

#include "hex8_le_core.cuh"

using namespace b200fem;

template <typename T>
static void apply(int L, ptrdiff_t ne, const int32_t *const *elements, const float *adj, const float *det, double mu,
                  double lambda, const T *u, T *out) {
    const int N = L + 1, NXE = N * N * N;
    std::vector<T> tu(3 * NXE), to(3 * NXE);
    for (ptrdiff_t e = 0; e < ne; ++e) {
        for (int n = 0; n < NXE; ++n)
            for (int c = 0; c < 3; ++c) {
                tu[c * NXE + n] = u[(ptrdiff_t)elements[n][e] * 3 + c];
                to[c * NXE + n] = 0;
            }
        T A[9];
        for (int k = 0; k < 9; ++k) A[k] = (T)adj[e * 9 + k];
        Hex8LEConsts<T> K;
        K.init(A, (T)det[e], T(1) / T(L), (T)mu, (T)lambda);
        for (int z = 0; z < L; ++z)
            for (int y = 0; y < L; ++y)
                for (int x = 0; x < L; ++x) {
                    T ul[3][8], ol[3][8];
                    for (int c = 0; c < 3; ++c)
                        for (int i = 0; i < 8; ++i)
                            ul[c][i] = tu[c * NXE + (z + ((i >> 2) & 1)) * N * N + (y + ((i >> 1) & 1)) * N + x + (i & 1)];
                    hex8_le_local_apply(K, ul, ol);
                    for (int c = 0; c < 3; ++c)
                        for (int i = 0; i < 8; ++i)
                            to[c * NXE + (z + ((i >> 2) & 1)) * N * N + (y + ((i >> 1) & 1)) * N + x + (i & 1)] += ol[c][i];
                }
        for (int n = 0; n < NXE; ++n)
            for (int c = 0; c < 3; ++c) out[(ptrdiff_t)elements[n][e] * 3 + c] += to[c * NXE + n];
    }
}

template <typename T>
static void block_diag(int L, ptrdiff_t ne, const int32_t *const *elements, const float *adj, const float *det, double mu,
                       double lambda, T *out6) {
    const int N = L + 1;
    const T h = T(1) / T(L);
    for (ptrdiff_t e = 0; e < ne; ++e) {
        T A[9];
        for (int k = 0; k < 9; ++k) A[k] = (T)adj[e * 9 + k] * (h * h);
        const T d = (T)det[e] * (h * h * h);
        for (int z = 0; z < N; ++z)
            for (int y = 0; y < N; ++y)
                for (int x = 0; x < N; ++x) {
                    T blk[6];
                    sshex8_le_node_block(A, d, (T)mu, (T)lambda, L, x, y, z, blk);
                    const ptrdiff_t id = elements[z * N * N + y * N + x][e];
                    for (int k = 0; k < 6; ++k) out6[id * 6 + k] += blk[k];
                }
    }
}

extern "C" int sshex8_le_core_apply_f64(int L, ptrdiff_t ne, const int32_t *const *el, const float *adj, const float *det,
                                        double mu, double lambda, const double *u, double *out) {
    apply<double>(L, ne, el, adj, det, mu, lambda, u, out);
    return 0;
}
extern "C" int sshex8_le_core_apply_f32(int L, ptrdiff_t ne, const int32_t *const *el, const float *adj, const float *det,
                                        double mu, double lambda, const float *u, float *out) {
    apply<float>(L, ne, el, adj, det, mu, lambda, u, out);
    return 0;
}
extern "C" int sshex8_le_core_block_diag_f64(int L, ptrdiff_t ne, const int32_t *const *el, const float *adj,
                                             const float *det, double mu, double lambda, double *out6) {
    block_diag<double>(L, ne, el, adj, det, mu, lambda, out6);
    return 0;
}
