// Host side of the warp-specialised sshex8 Laplacian kernel (sshex8_laplacian_ws.cuh): tensor map + launch.
#include <algorithm>

#include "sshex8_laplacian_ws.cuh"

namespace b200fem {

// 1-D tensor map over the input vector (box = WsLayout<T>::BOX elements): what the interior-plane copies go through.
// Encoded per call (the vector changes from call to call; ~1 us on the host).
template <typename T>
static int encode_vector_map(const T *x, ptrdiff_t nn, CUtensorMap *map) {
    typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                 const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                 CUtensorMapFloatOOBfill);
    static EncodeFn encode = [] {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
            fn = nullptr;
        return (EncodeFn)fn;
    }();
    if (!encode) B2_FAIL("cuTensorMapEncodeTiled is not available from this driver");
    const cuuint64_t dims[1] = {(cuuint64_t)nn};
    const cuuint64_t strides[1] = {0};  // rank 1: no strides are read
    const cuuint32_t box[1] = {(cuuint32_t)WsLayout<T>::BOX};
    const cuuint32_t estr[1] = {1};
    const CUresult r = encode(map, sizeof(T) == 8 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 1, const_cast<T *>(x), dims,
                              strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) B2_FAIL("cuTensorMapEncodeTiled failed");
    return B200FEM_SUCCESS;
}

template <typename T, int NC, int D, int S, bool PLAIN, bool PHASED = false>
static int launch_cfg(ptrdiff_t ne, ptrdiff_t nn, const Sshex8Header *headers, const T *x, T *y, int max_ctas, cudaStream_t s, int dbg) {
    constexpr int L = 8;
    using Cfg = Sshex8WsCfg<T, L, NC, D, S>;
    auto kern = sshex8_laplacian_ws_kernel<T, L, NC, D, S, PLAIN, PHASED>;
    static SmemOptIn opt;
    if (opt.ensure(kern, Cfg::smem_bytes)) return B200FEM_FAILURE;
    const ptrdiff_t ngroups = (ne + Cfg::EPT - 1) / Cfg::EPT;  // the last group may be partial (padded headers)
    if (ngroups == 0) return B200FEM_SUCCESS;
    alignas(64) CUtensorMap map;
    if (encode_vector_map(x, nn, &map)) return B200FEM_FAILURE;
    const int grid = (int)std::min<ptrdiff_t>((ngroups + NC - 1) / NC, (ptrdiff_t)max_ctas);
    kern<<<grid, Cfg::threads, Cfg::smem_bytes, s>>>(ne, headers, map, x, y, dbg);
    B2_CHECK_LAUNCH();
    return B200FEM_SUCCESS;
}

template <typename T, int NC, int D, bool PLAIN, bool PHASED = false, int DBG = 0, int FACEOUT = 0, int HINT = 0>
static int launch_ws2(ptrdiff_t ne, ptrdiff_t nn, const Sshex8Header *headers, const T *x, T *y, int max_ctas, cudaStream_t s, int dbg,
                      const WsDependency *dep = nullptr) {
    constexpr int L = 8;
    using Cfg = Sshex8WsCfg<T, L, NC, D, D, FACEOUT>;
    auto kern = sshex8_laplacian_ws2_kernel<T, L, NC, D, PLAIN, PHASED, DBG, FACEOUT, HINT>;
    static SmemOptIn opt;
    if (opt.ensure(kern, Cfg::smem_bytes)) return B200FEM_FAILURE;
    const ptrdiff_t ngroups = (ne + Cfg::EPT - 1) / Cfg::EPT;
    if (ngroups == 0) return B200FEM_SUCCESS;
    alignas(64) CUtensorMap map;
    if (encode_vector_map(x, nn, &map)) return B200FEM_FAILURE;
    const int grid = (int)std::min<ptrdiff_t>((ngroups + NC - 1) / NC, (ptrdiff_t)max_ctas);
    kern<<<grid, Cfg::threads, Cfg::smem_bytes, s>>>(ne, headers, map, x, y, dep ? *dep : WsDependency{});
    B2_CHECK_LAUNCH();
    return B200FEM_SUCCESS;
}

template <typename T, bool PLAIN>
static int launch_t(int cfg, ptrdiff_t ne, ptrdiff_t nn, const Sshex8Header *headers, const T *x, T *y, int max_ctas, cudaStream_t s,
                    const WsDependency *dep) {
    if (dep && (cfg & 255) != 11) B2_FAIL("the in-kernel ghost wait is built into ws_cfg 11 only");
    const int dbg = cfg >> 8;  // A/B switch (ws_cfg + 256): interior runs per node instead of TMA
    switch (cfg & 255) {
        // the round-1 kernel (four hand-offs per plane), kept for A/B
        case 0: return launch_cfg<T, 8, 4, 2, PLAIN>(ne, nn, headers, x, y, max_ctas, s, dbg);
        case 4: return launch_cfg<T, 8, 4, 2, PLAIN, true>(ne, nn, headers, x, y, max_ctas, s, dbg);  // phase-unrolled z loop, five bodies
        // the two-barrier kernel
        case 9: return launch_ws2<T, 8, 3, PLAIN>(ne, nn, headers, x, y, max_ctas, s, dbg);              // rolled z loop
        default:
        case 11: return launch_ws2<T, 8, 3, PLAIN, true>(ne, nn, headers, x, y, max_ctas, s, dbg, dep);  // + static accumulator phases: the default
        case 12: return launch_ws2<T, 8, 4, PLAIN, true>(ne, nn, headers, x, y, max_ctas, s, dbg);       // ring depth 4
        case 16: return launch_ws2<T, 8, 3, PLAIN, true, 0, 1>(ne, nn, headers, x, y, max_ctas, s, dbg);  // + lateral output buffer, all 32 positions
        case 18: return launch_ws2<T, 8, 3, PLAIN, true, 0, 2>(ne, nn, headers, x, y, max_ctas, s, dbg);  // + lateral output buffer, x = 0 / L only
        // bracketing runs of the default kernel (profiles/r02*_brackets*): parts switched off at compile time
        case 20: return launch_ws2<T, 8, 3, PLAIN, true, 2>(ne, nn, headers, x, y, max_ctas, s, dbg);          // no arithmetic
        case 21: return launch_ws2<T, 8, 3, PLAIN, true, 4 + 8>(ne, nn, headers, x, y, max_ctas, s, dbg);      // arithmetic only
        case 22: return launch_ws2<T, 8, 3, PLAIN, true, 2 + 4 + 8>(ne, nn, headers, x, y, max_ctas, s, dbg);  // protocol only
        case 23: return launch_ws2<T, 8, 3, PLAIN, true, 2 + 4>(ne, nn, headers, x, y, max_ctas, s, dbg);      // loads only
        case 24: return launch_ws2<T, 8, 3, PLAIN, true, 2 + 8>(ne, nn, headers, x, y, max_ctas, s, dbg);      // stores only
    }
}

int sshex8_laplacian_ws_launch(int real_type, bool plain, int cfg, ptrdiff_t ne, ptrdiff_t nn, const Sshex8Header *headers, const void *x,
                               void *y, int max_ctas, cudaStream_t s, const WsDependency *dep) {
    if (real_type == B200FEM_F64)
        return plain ? launch_t<double, true>(cfg, ne, nn, headers, (const double *)x, (double *)y, max_ctas, s, dep)
                     : launch_t<double, false>(cfg, ne, nn, headers, (const double *)x, (double *)y, max_ctas, s, dep);
    return plain ? launch_t<float, true>(cfg, ne, nn, headers, (const float *)x, (float *)y, max_ctas, s, dep)
                 : launch_t<float, false>(cfg, ne, nn, headers, (const float *)x, (float *)y, max_ctas, s, dep);
}

}  // namespace b200fem
