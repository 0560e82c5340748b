// interf.cu -- K2, interference power (placeholder until the tensor-core contraction lands).
#include "host_common.h"
using namespace wofdm;
extern "C" {
int wofdm_interf_power(wofdm_handle h, const wofdm_sys_t*, const double*, const double*, const double*, int, int, int, double*) {
    return fail(h, WOFDM_EUNSUPPORTED, "wofdm_interf_power: not built yet");
}
int wofdm_interf_power_scalar(wofdm_handle h, const wofdm_sys_t*, const double*, const double*, const double*, int, int, int, double*) {
    return fail(h, WOFDM_EUNSUPPORTED, "wofdm_interf_power_scalar: not built yet");
}
}
