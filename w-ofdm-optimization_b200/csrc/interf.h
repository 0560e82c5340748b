// interf.h -- shared between interf.cu (builders, fp64 path) and interf_tf32.cu (tcgen05 path).
#pragma once
#include "host_common.h"

namespace wofdm {

struct InterfDev {
    int N, n_tx, n_rx, N0, Kp, M, batch;
    double *vtx, *vrx;     // windows
    double2* chan;         // [C][L]
    double2* T;            // Tx_mat [n_tx][N]
    float2* T32;           // mode 1: the same, rounded to fp32 (band product of the TF32 path)
    double* Rbig;          // [2N][Kp], real form of Rx_mat, K interleaved (Re, Im) like Bbig's rows
    double* Bbig;          // [batch*Ms][Kp][N], rows 2b / 2b+1 = Re / Im B[b] per slice
    double* P;             // [C][N] (or [C] in scalar mode)
    float* tf32_work;      // mode 1: Rhi, Rlo [2N][Kp] and Bhi, Blo [batch*M][N][Kp] (K-major hi/lo splits)
};

// TF32-split tensor-core contraction of one batch (interf_tf32.cu)
// layout of the TF32 path's pre-tiled B operand (interf_tf32.cu): (column tile of TF32_TN sub-carriers j, K block of
// TF32_KB) = one contiguous [hi | lo] pair of K-major no-swizzle UMMA tiles; element (j, kk) of slice s, nk = Kp / TF32_KB,
// N / TF32_TN column tiles per slice (tile index sv = s * (N / TF32_TN) + j / TF32_TN)
constexpr int TF32_TN = 256, TF32_KB = 32;
__host__ __device__ inline size_t tf32_b_offset(int s, int nk, int N, int j, int kk) {
    const int sv = s * (N / TF32_TN) + j / TF32_TN, jj = j % TF32_TN;
    const int kb = kk / TF32_KB, k = kk % TF32_KB;
    const int c = (jj >> 3) * (TF32_KB / 4 * 8) + (k >> 2) * 8 + (jj & 7);    // 16-byte chunk inside the tile
    return ((size_t)sv * nk + kb) * (2 * (size_t)TF32_TN * TF32_KB) + (size_t)c * 4 + (k & 3);
}
// hi = fp32(x) with the 13 low mantissa bits cleared (a TF32 number), lo = fp32(x - hi)
__host__ __device__ inline void tf32_split(double x, float& hi, float& lo) {
    union { float f; unsigned u; } v;
    v.f = (float)x;
    v.u &= 0xffffe000u;
    hi = v.f;
    lo = (float)(x - (double)hi);
}

// k_isi: rows of K that exist for an ISI slice (interf.cu: interf_isi_k); slice 0 uses all Kp
// build_b<true> has written the hi/lo tiles of B into the work buffer (interf_tf32_b_tiles)
int interf_gemm_tf32(wofdm_ctx* h, const wofdm_sys_t* sys, const InterfDev& v, int Ms, int c0, int slices, int scalar, int k_isi);
float* interf_tf32_b_tiles(const InterfDev& v, int N);

}  // namespace wofdm
