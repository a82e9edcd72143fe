// Compressed connectivity of semi-structured hex8 meshes: 27 affine "entity boxes" per macro element
// (8 corners, 12 edges, 6 faces, 1 interior), id(x,y,z) = base + sx x + sy y + sz z inside a box. Shared by
// the sshex8 Laplacian, linear-elasticity and transfer plans (host-side fitting + the device descriptor).
#pragma once
#include <stdint.h>

#include <utility>
#include <vector>

namespace b200fem {

struct BoxDesc {
    int32_t base, sx, sy, sz;  // id = base + sx*x + sy*y + sz*z (absolute lattice coordinates)
};

// box class of a lattice position (cx + 3 cy + 9 cz, c = 0 low face, 1 inside, 2 high face)
__host__ __device__ __forceinline__ int sshex8_box_class(int L, int x, int y, int z) {
    const int cx = (x == 0) ? 0 : (x == L ? 2 : 1);
    const int cy = (y == 0) ? 0 : (y == L ? 2 : 1);
    const int cz = (z == 0) ? 0 : (z == L ? 2 : 1);
    return cx + 3 * cy + 9 * cz;
}

// Fit the 27 entity boxes of every element; returns false if some box is not affine in the lattice.
// Also decides whether the interior nodes are exclusive to their element (compact (L-1)^3 blocks that
// no other element references). If so, `zero_ranges` receives the complementary id ranges -- the
// skeleton, everything that receives RED.ADD contributions -- which is what an overwriting apply has to
// zero first (one range for the reference's numbering, two for an [owned | ghost] local numbering).
inline bool fit_boxes(int L, ptrdiff_t ne, ptrdiff_t nn, const int32_t *const *elements, std::vector<BoxDesc> &out,
                      bool *exclusive, std::vector<std::pair<int64_t, int64_t>> &zero_ranges) {
    const int N = L + 1;
    out.resize((size_t)ne * 27);
    auto lidx = [N](int x, int y, int z) { return z * N * N + y * N + x; };
    const int lo[3] = {0, 1, L}, hi[3] = {0, L - 1, L};
    bool compact = true;
    for (ptrdiff_t e = 0; e < ne; ++e) {
        for (int cz = 0; cz < 3; ++cz)
            for (int cy = 0; cy < 3; ++cy)
                for (int cx = 0; cx < 3; ++cx) {
                    const int x0 = lo[cx], y0 = lo[cy], z0 = lo[cz];
                    const int x1 = hi[cx], y1 = hi[cy], z1 = hi[cz];
                    const int64_t b0 = elements[lidx(x0, y0, z0)][e];
                    const int64_t sx = (x1 > x0) ? (int64_t)elements[lidx(x0 + 1, y0, z0)][e] - b0 : 0;
                    const int64_t sy = (y1 > y0) ? (int64_t)elements[lidx(x0, y0 + 1, z0)][e] - b0 : 0;
                    const int64_t sz = (z1 > z0) ? (int64_t)elements[lidx(x0, y0, z0 + 1)][e] - b0 : 0;
                    for (int z = z0; z <= z1; ++z)
                        for (int y = y0; y <= y1; ++y)
                            for (int x = x0; x <= x1; ++x) {
                                const int64_t id = elements[lidx(x, y, z)][e];
                                if (id != b0 + sx * (x - x0) + sy * (y - y0) + sz * (z - z0)) return false;
                            }
                    const int64_t base = b0 - sx * x0 - sy * y0 - sz * z0;
                    if (base < INT32_MIN || base > INT32_MAX) return false;
                    out[(size_t)e * 27 + cx + 3 * cy + 9 * cz] = BoxDesc{(int32_t)base, (int32_t)sx, (int32_t)sy, (int32_t)sz};
                    if (cx == 1 && cy == 1 && cz == 1 && L > 2) {
                        if (sx != 1 || sy != (L - 1) || sz != (int64_t)(L - 1) * (L - 1)) compact = false;
                    }
                }
    }
    zero_ranges.clear();
    bool excl = compact;
    if (excl) {
        // reference counts: interior ids must be hit exactly once overall, by an interior position
        std::vector<uint8_t> interior((size_t)nn, 0);
        const int64_t vol = (int64_t)(L - 1) * (L - 1) * (L - 1);
        for (ptrdiff_t e = 0; e < ne && excl; ++e) {
            const int64_t b = elements[lidx(1, 1, 1)][e];
            if (b + vol > nn) excl = false;
            for (int64_t k = 0; k < vol && excl; ++k) {
                if (interior[(size_t)(b + k)]) excl = false;
                interior[(size_t)(b + k)] = 1;
            }
        }
        for (ptrdiff_t e = 0; e < ne && excl; ++e)
            for (int z = 0; z <= L && excl; ++z)
                for (int y = 0; y <= L && excl; ++y)
                    for (int x = 0; x <= L; ++x) {
                        const bool in = x > 0 && x < L && y > 0 && y < L && z > 0 && z < L;
                        if (!in && interior[(size_t)elements[lidx(x, y, z)][e]]) {
                            excl = false;
                            break;
                        }
                    }
        if (excl) {
            int64_t start = -1;
            for (int64_t i = 0; i <= (int64_t)nn; ++i) {
                const bool skel = (i < (int64_t)nn) && !interior[(size_t)i];
                if (skel && start < 0) start = i;
                if (!skel && start >= 0) {
                    zero_ranges.emplace_back(start, i);
                    start = -1;
                }
            }
            if (zero_ranges.size() > 8) {  // scattered skeleton: not worth a memset per run
                zero_ranges.clear();
                excl = false;
            }
        }
    }
    *exclusive = excl;
    return true;
}


}  // namespace b200fem
