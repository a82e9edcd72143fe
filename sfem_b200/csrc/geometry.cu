// Geometry caches on the device: FFF (Laplacian) and adjugate/determinant (elasticity).
//
// Replace tet4_fff_fill (operators/tet4/tet4_fff.cpp:4-27), hex8_fff_fill
// (operators/hex8/hex8_fff.cpp:5-37, evaluated at the reference point (1/2,1/2,1/2): affine
// assumption) and hex8_adjugate_and_det_fill (operators/hex8/hex8_jacobian.cpp:5-38).
// Restated from the definitions: J[r][c] = d x_r / d xi_c, adj = det(J) J^-1 (row j = det * grad xi_j),
//   tet4: FFF = adj adj^T / (6 det)     hex8: FFF = adj adj^T / det .
// Arithmetic is fp64 from the fp32 coordinates, results stored fp32 (jacobian_t). The reference's
// tet4 path evaluates in fp32; the two agree to fp32 rounding (tests/test_gpu_operators.py::test_geometry_fill), and operator
// parity tests always hand both sides the same fp32 arrays.
#include <vector>

#include "common.cuh"

namespace b200fem {

__device__ __forceinline__ void adjugate3(const double J[9], double A[9], double &det) {
    A[0] = J[4] * J[8] - J[5] * J[7];
    A[1] = J[2] * J[7] - J[1] * J[8];
    A[2] = J[1] * J[5] - J[2] * J[4];
    A[3] = J[5] * J[6] - J[3] * J[8];
    A[4] = J[0] * J[8] - J[2] * J[6];
    A[5] = J[2] * J[3] - J[0] * J[5];
    A[6] = J[3] * J[7] - J[4] * J[6];
    A[7] = J[1] * J[6] - J[0] * J[7];
    A[8] = J[0] * J[4] - J[1] * J[3];
    det = J[0] * A[0] + J[1] * A[3] + J[2] * A[6];
}

__device__ __forceinline__ void hex8_center_jacobian(const ptrdiff_t e, const ptrdiff_t ne, const int32_t *elements,
                                                     const float *px, const float *py, const float *pz, double J[9]) {
    // VTK corner signs of d/dxi, d/deta, d/dzeta at the centre (each times 1/4)
    const int sx[8] = {-1, 1, 1, -1, -1, 1, 1, -1};
    const int sy[8] = {-1, -1, 1, 1, -1, -1, 1, 1};
    const int sz[8] = {-1, -1, -1, -1, 1, 1, 1, 1};
#pragma unroll
    for (int k = 0; k < 9; ++k) J[k] = 0;
#pragma unroll
    for (int v = 0; v < 8; ++v) {
        const int32_t i = elements[v * ne + e];
        const double p[3] = {(double)px[i], (double)py[i], (double)pz[i]};
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            J[r * 3 + 0] += 0.25 * sx[v] * p[r];
            J[r * 3 + 1] += 0.25 * sy[v] * p[r];
            J[r * 3 + 2] += 0.25 * sz[v] * p[r];
        }
    }
}

__device__ __forceinline__ void fff_from_adjugate(const double A[9], const double scale, float *fff) {
    int k = 0;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = i; j < 3; ++j)
            fff[k++] = (float)((A[i * 3 + 0] * A[j * 3 + 0] + A[i * 3 + 1] * A[j * 3 + 1] + A[i * 3 + 2] * A[j * 3 + 2]) * scale);
}

__global__ void tet4_fff_kernel(const ptrdiff_t ne, const int32_t *__restrict__ elements, const float *__restrict__ px,
                                const float *__restrict__ py, const float *__restrict__ pz, float *__restrict__ fff) {
    for (ptrdiff_t e = (ptrdiff_t)blockIdx.x * blockDim.x + threadIdx.x; e < ne; e += (ptrdiff_t)gridDim.x * blockDim.x) {
        int32_t ev[4];
#pragma unroll
        for (int v = 0; v < 4; ++v) ev[v] = elements[v * ne + e];
        double J[9];
        const float *P[3] = {px, py, pz};
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const double p0 = P[r][ev[0]];
#pragma unroll
            for (int c = 0; c < 3; ++c) J[r * 3 + c] = (double)P[r][ev[c + 1]] - p0;
        }
        double A[9], det;
        adjugate3(J, A, det);
        fff_from_adjugate(A, 1.0 / (6.0 * det), &fff[e * 6]);
    }
}

__global__ void hex8_fff_kernel(const ptrdiff_t ne, const int32_t *__restrict__ elements, const float *__restrict__ px,
                                const float *__restrict__ py, const float *__restrict__ pz, float *__restrict__ fff) {
    for (ptrdiff_t e = (ptrdiff_t)blockIdx.x * blockDim.x + threadIdx.x; e < ne; e += (ptrdiff_t)gridDim.x * blockDim.x) {
        double J[9], A[9], det;
        hex8_center_jacobian(e, ne, elements, px, py, pz, J);
        adjugate3(J, A, det);
        fff_from_adjugate(A, 1.0 / det, &fff[e * 6]);
    }
}

__global__ void hex8_adjugate_kernel(const ptrdiff_t ne, const int32_t *__restrict__ elements,
                                     const float *__restrict__ px, const float *__restrict__ py,
                                     const float *__restrict__ pz, float *__restrict__ adjugate,
                                     float *__restrict__ determinant) {
    for (ptrdiff_t e = (ptrdiff_t)blockIdx.x * blockDim.x + threadIdx.x; e < ne; e += (ptrdiff_t)gridDim.x * blockDim.x) {
        double J[9], A[9], det;
        hex8_center_jacobian(e, ne, elements, px, py, pz, J);
        adjugate3(J, A, det);
#pragma unroll
        for (int k = 0; k < 9; ++k) adjugate[e * 9 + k] = (float)A[k];
        determinant[e] = (float)det;
    }
}

// tet4_adjugate_and_det (operators/tet4/tet4_inline_cpu.hpp:285-380): J[r][c] = x_r(v_{c+1}) - x_r(v_0), adjugate = det J^-1
__global__ void tet4_adjugate_kernel(const ptrdiff_t ne, const int32_t *__restrict__ elements, const float *__restrict__ px,
                                     const float *__restrict__ py, const float *__restrict__ pz, float *__restrict__ adjugate,
                                     float *__restrict__ determinant) {
    for (ptrdiff_t e = (ptrdiff_t)blockIdx.x * blockDim.x + threadIdx.x; e < ne; e += (ptrdiff_t)gridDim.x * blockDim.x) {
        int32_t ev[4];
#pragma unroll
        for (int v = 0; v < 4; ++v) ev[v] = elements[v * ne + e];
        double J[9];
        const float *P[3] = {px, py, pz};
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const double p0 = P[r][ev[0]];
#pragma unroll
            for (int c = 0; c < 3; ++c) J[r * 3 + c] = (double)P[r][ev[c + 1]] - p0;
        }
        double A[9], det;
        adjugate3(J, A, det);
#pragma unroll
        for (int k = 0; k < 9; ++k) adjugate[e * 9 + k] = (float)A[k];
        determinant[e] = (float)det;
    }
}

struct GeomInputs {
    DeviceArray<int32_t> elements;
    DeviceArray<float> points;  // [3][nn]
    int upload(int nxe, ptrdiff_t ne, ptrdiff_t nn, const int32_t *const *el, const float *const *pts) {
        std::vector<int32_t> h((size_t)nxe * ne);
        for (int v = 0; v < nxe; ++v) {
            for (ptrdiff_t e = 0; e < ne; ++e) {
                if (el[v][e] < 0 || el[v][e] >= nn) B2_FAIL("geometry fill: node index out of range");
                h[(size_t)v * ne + e] = el[v][e];
            }
        }
        if (elements.upload(h.data(), h.size())) return B200FEM_FAILURE;
        if (points.alloc((size_t)3 * nn)) return B200FEM_FAILURE;
        for (int d = 0; d < 3; ++d)
            B2_CUDA(cudaMemcpy(points.ptr + (size_t)d * nn, pts[d], (size_t)nn * sizeof(float), cudaMemcpyHostToDevice));
        return B200FEM_SUCCESS;
    }
};

}  // namespace b200fem

using namespace b200fem;

// the 8 macro corners inside an sshex8 connectivity are not known here: callers pass them as `elements`
static int corner_count(int element_type) {
    switch (element_type) {
        case B200FEM_TET4:
        case B200FEM_TET10:       // 10-node tetrahedra: the affine map of the 4 corners (first 4 rows of `elements`), as
        case B200FEM_MACRO_TET4:  // tet10_laplacian_apply / macro_tet4_laplacian_apply do (tet4_fff_s on ev[0..3])
            return 4;
        case B200FEM_HEX8:
        case B200FEM_SSHEX8: return 8;
        default: return 0;
    }
}

extern "C" {

int b200fem_fff_fill(int element_type, ptrdiff_t ne, ptrdiff_t nn, const int32_t *const *elements,
                     const float *const *points, float *fff) {
    const int nxe = corner_count(element_type);
    if (!nxe) B2_FAIL("fff_fill: unsupported element type");
    if (ne == 0) return B200FEM_SUCCESS;
    if (!elements || !points || !fff) B2_FAIL("fff_fill: NULL argument");
    GeomInputs in;
    if (in.upload(nxe, ne, nn, elements, points)) return B200FEM_FAILURE;
    DeviceArray<float> out;
    if (out.alloc((size_t)ne * 6)) return B200FEM_FAILURE;
    const int grid = (int)std::min<ptrdiff_t>((ne + 255) / 256, (ptrdiff_t)sm_count() * 16);
    const float *px = in.points.ptr, *py = px + nn, *pz = py + nn;
    if (nxe == 4)
        tet4_fff_kernel<<<grid, 256>>>(ne, in.elements.ptr, px, py, pz, out.ptr);
    else
        hex8_fff_kernel<<<grid, 256>>>(ne, in.elements.ptr, px, py, pz, out.ptr);
    B2_CHECK_LAUNCH();
    B2_CUDA(cudaMemcpy(fff, out.ptr, out.bytes(), cudaMemcpyDeviceToHost));
    return B200FEM_SUCCESS;
}

int b200fem_adjugate_and_det_fill(int element_type, ptrdiff_t ne, ptrdiff_t nn, const int32_t *const *elements,
                                  const float *const *points, float *adjugate, float *determinant) {
    const int nxe = corner_count(element_type);
    if (!nxe) B2_FAIL("adjugate_and_det_fill: tet4 / tet10 / macro_tet4 / hex8 / sshex8 only");
    if (ne == 0) return B200FEM_SUCCESS;
    if (!elements || !points || !adjugate || !determinant) B2_FAIL("adjugate_and_det_fill: NULL argument");
    GeomInputs in;
    if (in.upload(nxe, ne, nn, elements, points)) return B200FEM_FAILURE;
    DeviceArray<float> adj, det;
    if (adj.alloc((size_t)ne * 9) || det.alloc((size_t)ne)) return B200FEM_FAILURE;
    const int grid = (int)std::min<ptrdiff_t>((ne + 255) / 256, (ptrdiff_t)sm_count() * 16);
    const float *px = in.points.ptr, *py = px + nn, *pz = py + nn;
    if (nxe == 4)
        tet4_adjugate_kernel<<<grid, 256>>>(ne, in.elements.ptr, px, py, pz, adj.ptr, det.ptr);
    else
        hex8_adjugate_kernel<<<grid, 256>>>(ne, in.elements.ptr, px, py, pz, adj.ptr, det.ptr);
    B2_CHECK_LAUNCH();
    B2_CUDA(cudaMemcpy(adjugate, adj.ptr, adj.bytes(), cudaMemcpyDeviceToHost));
    B2_CUDA(cudaMemcpy(determinant, det.ptr, det.bytes(), cudaMemcpyDeviceToHost));
    return B200FEM_SUCCESS;
}

}  // extern "C"
