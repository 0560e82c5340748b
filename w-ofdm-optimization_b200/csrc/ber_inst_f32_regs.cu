// fp32 instantiations of the register-resident K1 policy (production path).
// WOFDM_VARIANT(type, N, threads, TC = samples per thread, LB = taps in registers (L <= LB), min CTAs/SM,
//               FULL = threads*TC == S*stride exactly, tag)
#include "ber_registry.h"
namespace wofdm {
void register_ber_f32_regs(std::vector<BerVariant>& out) {
    WOFDM_VARIANT(float, 256, 256, 17, 21, 2, true, "f32r")
    WOFDM_VARIANT(float, 256, 256, 17, 21, 2, false, "f32r")
    WOFDM_VARIANT(float, 256, 256, 19, 21, 2, false, "f32r")
    WOFDM_VARIANT_CIRC(float, 256, 256, 17, 21, 2, true, "f32r")
    WOFDM_VARIANT_CIRC(float, 256, 256, 17, 21, 2, false, "f32r")
    WOFDM_VARIANT_CIRC(float, 256, 256, 19, 21, 2, false, "f32r")
    // N = 1024: one frame per 2-CTA cluster, 8 OFDM symbols and 512 threads per CTA (ber_kernel.cuh, CL)
    WOFDM_VARIANT_CL(float, 1024, 512, 17, 21, 1, true, 2, "f32r")
    WOFDM_VARIANT_CL(float, 1024, 512, 17, 21, 1, false, 2, "f32r")
    WOFDM_VARIANT_CL(float, 1024, 512, 19, 21, 1, false, 2, "f32r")
}
}  // namespace wofdm
