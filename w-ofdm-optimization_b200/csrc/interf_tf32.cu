// interf_tf32.cu -- TF32-split (3xTF32) tcgen05 contraction for the interference power (mode 1).
#include "interf.h"

namespace wofdm {

int interf_gemm_tf32(wofdm_ctx* h, const wofdm_sys_t*, const InterfDev&, int, int, int, int) {
    return fail(h, WOFDM_EUNSUPPORTED, "TF32-split interference path is not built yet (use mode 0)");
}

}  // namespace wofdm
