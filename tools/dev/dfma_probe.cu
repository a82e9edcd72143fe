// Microbenchmark: DFMA issue rate per SM sub-partition on sm_100a as a function of warps per scheduler and
// independent accumulator chains per thread (what the sshex8 compute warps can expect from the FP64 pipe).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o dfma_probe dfma_probe.cu && ./dfma_probe
#include <cstdio>
#include <cuda_runtime.h>

template <int K, bool WITH_LDS>
__global__ void probe(double *out, long long *cycles, int iters, double a, double b) {
    __shared__ double sm[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = a * i;
    __syncthreads();
    double acc[K];
#pragma unroll
    for (int k = 0; k < K; ++k) acc[k] = threadIdx.x + k;
    double c0 = a, c1 = b, c2 = a + b;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        double r0 = c0, r1 = c1, r2 = c2;
        if (WITH_LDS) {
            r0 = sm[(threadIdx.x + it) & 1023];
            r1 = sm[(threadIdx.x + it + 9) & 1023];
            r2 = sm[(threadIdx.x + it + 18) & 1023];
        }
#pragma unroll
        for (int k = 0; k < K; ++k) acc[k] = __fma_rn(acc[k], r0, r1);
#pragma unroll
        for (int k = 0; k < K; ++k) acc[k] = __fma_rn(acc[k], r1, r2);
#pragma unroll
        for (int k = 0; k < K; ++k) acc[k] = __fma_rn(acc[k], r2, r0);
    }
    const long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int k = 0; k < K; ++k) s += acc[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int K, bool WITH_LDS>
void run(int warps_per_sm) {
    double *out;
    long long *cyc;
    cudaMalloc(&out, 148 * 1024 * sizeof(double));
    cudaMalloc(&cyc, 148 * sizeof(long long));
    const int iters = 20000;
    probe<K, WITH_LDS><<<148, warps_per_sm * 32>>>(out, cyc, iters, 1.0000001, 1e-9);
    cudaDeviceSynchronize();
    probe<K, WITH_LDS><<<148, warps_per_sm * 32>>>(out, cyc, iters, 1.0000001, 1e-9);
    cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    const double per_warp_dfma = (double)h[0] / ((double)iters * 3 * K);
    const double per_smsp = per_warp_dfma / (warps_per_sm / 4.0);
    printf("{\"chains\": %d, \"lds\": %d, \"warps_per_scheduler\": %.1f, \"cycles_per_dfma_per_warp\": %.3f, \"cycles_per_dfma_per_scheduler\": %.3f}\n", K,
           (int)WITH_LDS, warps_per_sm / 4.0, per_warp_dfma, per_smsp);
    cudaFree(out);
    cudaFree(cyc);
}

int main() {
    for (int w : {4, 8, 12, 16}) {
        run<1, false>(w);
        run<2, false>(w);
        run<4, false>(w);
        run<9, false>(w);
        run<27, false>(w);
        run<27, true>(w);
    }
    return 0;
}
