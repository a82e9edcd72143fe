// sshex8 linear elasticity (placeholder translation unit: the kernel is built in a later milestone).
#include "common.cuh"

namespace b200fem {

OpBase *make_sshex8_linear_elasticity(int, ptrdiff_t, ptrdiff_t, const int32_t *const *, const float *, const float *,
                                      double, double, int) {
    set_error("sshex8 linear elasticity: not built yet");
    return nullptr;
}

}  // namespace b200fem
