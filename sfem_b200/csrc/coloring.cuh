// Lazy element-colouring state shared by the thread-per-element operators.
#pragma once
#include <stdlib.h>

#include <algorithm>
#include <vector>

#include "common.cuh"

namespace b200fem {

struct ColorPlan {
    DeviceArray<int32_t> perm;           // elements grouped by colour
    std::vector<int32_t> offsets;        // [ncolors+1]
    std::vector<std::vector<int32_t>> host_elements;  // SoA copy kept for the lazy build
    bool ready = false;

    void keep(int nxe, ptrdiff_t ne, const int32_t *const *elements) {
        host_elements.resize(nxe);
        for (int v = 0; v < nxe; ++v) host_elements[v].assign(elements[v], elements[v] + ne);
    }

    int build(ptrdiff_t ne, ptrdiff_t nn) {
        if (ready) return B200FEM_SUCCESS;
        const int nxe = (int)host_elements.size();
        std::vector<const int32_t *> ptrs(nxe);
        for (int v = 0; v < nxe; ++v) ptrs[v] = host_elements[v].data();
        int32_t *p = nullptr, *off = nullptr;
        int nc = 0;
        std::string err;
        color_elements(ne, nn, nxe, ptrs.data(), &err, &p, &off, &nc);
        if (!p) B2_FAIL(err);
        offsets.assign(off, off + nc + 1);
        const int rc = perm.upload(p, (size_t)ne);
        free(p);
        free(off);
        if (rc) return rc;
        ready = true;
        host_elements.clear();
        host_elements.shrink_to_fit();
        return B200FEM_SUCCESS;
    }
    int ncolors() const { return ready ? (int)offsets.size() - 1 : 0; }
};

static inline bool check_indices(int nxe, ptrdiff_t ne, ptrdiff_t nn, const int32_t *const *elements) {
    for (int v = 0; v < nxe; ++v) {
        const int32_t *row = elements[v];
        for (ptrdiff_t e = 0; e < ne; ++e)
            if (row[e] < 0 || row[e] >= nn) return false;
    }
    return true;
}

}  // namespace b200fem
