// mask_kernel.cuh -- Tx side of the channel-mask BER variant (matlab/main_channel_mask.m).
//
// One CTA builds the serialised, MASKED Tx stream of one frame and writes it to HBM for the staged K1 kernel
// (BerParams::tx_stream): the same Philox symbol draws as K1 (so Rx can redraw them), guard band, register IFFT,
// CP/CS + Tx window, then the DFT-domain raised-cosine mask of gen_tx_ofdm / dft_rc_filt
// (main_channel_mask.m:384-417): per symbol the zero-padded windowed symbol is circularly convolved (mod
// M = 2 n_tx - 1) with g = IDFT_M(ifftshift(windowRC)); samples 0..n_tx-1 stay in the symbol, the other n_tx - 1 are
// added to the start of the NEXT symbol; finally the symbols are overlap-added with the frame stride (tx2rx, :420-431).
// Round 1 evaluates the convolution in its dense time-domain form from shared memory (n_tx complex MACs per output,
// 27x the flops of the rest of the chain): correct and on the device, ~2e7 OFDM symbols/s; the two-2048-point-FFT
// form sketched in DESIGN.md section 7 is the next step.  fp32.
#pragma once
#include "ber_kernel.cuh"

namespace wofdm {

struct MaskParams {
    int N, cp, cs, tail_tx, bits, S, n_tx, stride, constellation, guard;
    int M;                         // 2 n_tx - 1
    const float* win_tx;           // [n_tx], v_tx * qam_scale / N (the K1 table)
    const float2* tw;              // FFT twiddles (the K1 table)
    const float2* g;               // [M] impulse response of the mask
    unsigned long long seed;
    long long frame_begin, frame_step, n_frames;
    float2* stream;                // [n_frames][tail_tx + S*stride]
};

template <int N> struct MaskSmem {
    using P = FftPlan<N>;
    static constexpr int NT = 256, FPP = NT / P::TPF;
    static constexpr int RMAX = (3 * N + NT - 1) / NT;     // outputs per thread: M = 2 n_tx - 1 <= RMAX * NT (n_tx <= 1.5 N)
    static size_t bytes(int S, int n_tx, int stride, int tail_tx, int M) {
        const size_t body = (size_t)tail_tx + (size_t)S * stride;
        return ((size_t)S * N + (size_t)FPP * P::XLEN + P::NTW + 256 + n_tx + M + body + 8) * sizeof(float2);
    }
};

template <int N>
__global__ void __launch_bounds__(256) tx_mask_kernel(const MaskParams p) {
    using P = FftPlan<N>;
    constexpr int NT = 256, TPF = P::TPF, FPP = NT / TPF;
    extern __shared__ __align__(16) unsigned char msm[];
    const int tid = threadIdx.x, slot = tid / TPF, t = tid % TPF;
    const int S = p.S, n_tx = p.n_tx, stride = p.stride, M = p.M;
    const int body = p.tail_tx + S * stride;
    const int hb = p.bits >> 1, m = 1 << hb;
    float2* xs = reinterpret_cast<float2*>(msm);           // [S][N] un-windowed IFFT outputs
    float2* xb = xs + (size_t)S * N;                       // FFT exchange
    float2* tw = xb + (size_t)FPP * P::XLEN;
    float2* qlut = tw + P::NTW;
    float2* ts = qlut + 256;                               // [n_tx] one windowed symbol
    float2* gs = ts + n_tx;                                // [M]
    float2* us = gs + M;                                   // [body] the frame stream being accumulated
    for (int i = tid; i < P::NTW; i += NT) tw[i] = p.tw[i];
    for (int i = tid; i < M; i += NT) gs[i] = p.g[i];
    for (int i = tid; i < 256; i += NT) {
        float2 v = make_float2(0.f, 0.f);
        if (i < (1 << p.bits)) {
            int a, c;
            idx_to_levels(i, hb, m, p.constellation, a, c);
            v = make_float2((float)(2 * a - (m - 1)), (float)(2 * c - (m - 1)));
        }
        qlut[i] = v;
    }
    __syncthreads();
    BerParams draw = {};                                   // only what load_sym_idx reads
    draw.seed = p.seed; draw.bits = p.bits; draw.S = S;
    constexpr int RMAX = MaskSmem<N>::RMAX;
    for (long long j = blockIdx.x; j < p.n_frames; j += gridDim.x) {
        const long long f = p.frame_begin + j * p.frame_step;
        for (int i = tid; i < body; i += NT) us[i] = make_float2(0.f, 0.f);
        // ---- symbols -> IFFT (un-windowed, un-scaled), FPP symbols per pass
        for (int s0 = 0; s0 < S; s0 += FPP) {
            const int s = s0 + slot, se = s < S ? s : S - 1;
            uint32_t w[4];
            load_sym_idx<N, false>(draw, f, se, t, w);
            float2 v[16];
#pragma unroll
            for (int q = 0; q < 16; ++q)
                v[q] = bin_active<N>(t + q * TPF, p.guard) ? qlut[sym_byte(w, q)] : make_float2(0.f, 0.f);
            fft_regs<float, N, +1, FPP>(v, t, xb + slot * P::XLEN, tw, slot);
#pragma unroll
            for (int q = 0; q < 16; ++q) xs[(size_t)se * N + t + q * TPF] = v[q];
            __syncthreads();
        }
        // ---- per symbol: window + CP/CS, mask, accumulate into the stream
        for (int s = 0; s < S; ++s) {
            for (int i = tid; i < n_tx; i += NT) ts[i] = cscale(p.win_tx[i], xs[(size_t)s * N + ((i - p.cp) & (N - 1))]);
            __syncthreads();
            float2 y[RMAX];
#pragma unroll
            for (int r = 0; r < RMAX; ++r) {
                const int n = tid + r * NT;
                float2 a0 = make_float2(0.f, 0.f), a1 = make_float2(0.f, 0.f);
                if (n < M) {
                    // y[n] = sum_k ts[k] g[(n - k) mod M]: k <= n reads g[n-k], k > n reads g[M + n - k]
                    const int k1 = n < n_tx - 1 ? n : n_tx - 1;
                    int k = 0;
                    for (; k + 1 <= k1; k += 2) { cmac(a0, ts[k], gs[n - k]); cmac(a1, ts[k + 1], gs[n - k - 1]); }
                    for (; k <= k1; ++k) cmac(a0, ts[k], gs[n - k]);
                    for (; k < n_tx; ++k) cmac(a1, ts[k], gs[M + n - k]);
                }
                y[r] = cadd(a0, a1);
            }
            // samples 0..n_tx-1 stay in symbol s (positions s*stride + n) ...
#pragma unroll
            for (int r = 0; r < RMAX; ++r) {
                const int n = tid + r * NT;
                if (n < n_tx) us[s * stride + n] = cadd(us[s * stride + n], y[r]);
            }
            __syncthreads();
            // ... the filter's tail goes to the start of symbol s+1; the last symbol's tail is dropped (:413-416)
            if (s + 1 < S) {
#pragma unroll
                for (int r = 0; r < RMAX; ++r) {
                    const int n = tid + r * NT;
                    if (n >= n_tx && n < M) us[(s + 1) * stride + (n - n_tx)] = cadd(us[(s + 1) * stride + (n - n_tx)], y[r]);
                }
            }
            __syncthreads();
        }
        float2* dst = p.stream + (size_t)j * body;
        for (int i = tid; i < body; i += NT) dst[i] = us[i];
        __syncthreads();
    }
}

}  // namespace wofdm
