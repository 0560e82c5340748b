// interf.h -- shared between interf.cu (builders, fp64 path) and interf_tf32.cu (tcgen05 path).
#pragma once
#include "host_common.h"

namespace wofdm {

struct InterfDev {
    int N, n_tx, n_rx, N0, Kp, M, batch;
    double *vtx, *vrx;     // windows
    double2* chan;         // [C][L]
    double2* T;            // Tx_mat [n_tx][N]
    double* Rbig;          // [2N][Kp] = [Rr -Ri; Ri Rr]
    double* Bbig;          // [batch*Ms][Kp][N] = [Re B; Im B] per slice
    double* P;             // [C][N] (or [C] in scalar mode)
    float* tf32_work;      // mode 1: Rhi, Rlo [2N][Kp] and Bhi, Blo [batch*M][N][Kp] (K-major hi/lo splits)
};

// TF32-split tensor-core contraction of one batch (interf_tf32.cu)
int interf_gemm_tf32(wofdm_ctx* h, const wofdm_sys_t* sys, const InterfDev& v, int Ms, int c0, int slices, int scalar);

}  // namespace wofdm
