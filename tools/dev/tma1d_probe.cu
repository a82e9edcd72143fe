// Probe: does a 1-D tensor-map TMA load work for (dtype, box, coordinate)? One variant per process.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <typename T>
__global__ void probe(const __grid_constant__ CUtensorMap map, int coord, int box, T *out, int dst_off) {
    extern __shared__ __align__(128) unsigned char sm[];
    T *buf = reinterpret_cast<T *>(sm + 128) + dst_off;
    uint64_t *bar = reinterpret_cast<uint64_t *>(sm);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(s32(bar)));
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(s32(bar)), "r"((uint32_t)(box * sizeof(T))) : "memory");
        asm volatile("cp.async.bulk.tensor.1d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2}], [%3];\n" ::"r"(s32(buf)),
                     "l"(reinterpret_cast<uint64_t>(&map)), "r"(coord), "r"(s32(bar))
                     : "memory");
    }
    asm volatile("{\n .reg .pred p;\n W: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n @p bra Dn;\n bra W;\n Dn:\n}\n" ::"r"(s32(bar)) : "memory");
    for (int i = threadIdx.x; i < box; i += blockDim.x) out[i] = buf[i];
}
template <typename T>
int run(int box, int coord, int dst_off) {
    const int n = 4096;
    T *h = (T *)malloc(n * sizeof(T));
    for (int i = 0; i < n; ++i) h[i] = (T)i;
    T *d, *o;
    cudaMalloc(&d, n * sizeof(T));
    cudaMalloc(&o, 512 * sizeof(T));
    cudaMemcpy(d, h, n * sizeof(T), cudaMemcpyHostToDevice);
    typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                 const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    alignas(64) CUtensorMap map;
    const cuuint64_t dims[1] = {(cuuint64_t)n};
    const cuuint64_t strides[1] = {0};
    const cuuint32_t bx[1] = {(cuuint32_t)box};
    const cuuint32_t es[1] = {1};
    CUresult r = ((EncodeFn)fn)(&map, sizeof(T) == 8 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 1, d, dims, strides, bx, es,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 2; }
    probe<T><<<1, 32, 128 + 4096>>>(map, coord, box, o, dst_off);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(e)); return 1; }
    T res[512];
    cudaMemcpy(res, o, box * sizeof(T), cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int i = 0; i < box; ++i) {
        const int g = coord + i;
        const T want = (g >= 0 && g < n) ? (T)g : (T)0;
        if (res[i] != want) ++bad;
    }
    printf("ok, mismatches %d (first %g last %g)\n", bad, (double)res[0], (double)res[box - 1]);
    return 0;
}
int main(int argc, char **argv) {
    const int f64 = atoi(argv[1]), box = atoi(argv[2]), coord = atoi(argv[3]), dst_off = argc > 4 ? atoi(argv[4]) : 0;
    printf("f64=%d box=%d coord=%d dst_off=%d: ", f64, box, coord, dst_off);
    return f64 ? run<double>(box, coord, dst_off) : run<float>(box, coord, dst_off);
}
