// fp32 instantiations of the register-resident K1 policy (production path).
#include "ber_registry.h"
namespace wofdm {
void register_ber_f32_regs(std::vector<BerVariant>& out) {
    WOFDM_VARIANT(float, 256, 256, 17, 21, "f32r")
    WOFDM_VARIANT(float, 256, 256, 19, 21, "f32r")
    WOFDM_VARIANT(float, 1024, 512, 35, 11, "f32r")
}
}  // namespace wofdm
