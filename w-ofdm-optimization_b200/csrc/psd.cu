// psd.cu -- next row 8f-3: PSD / out-of-band-radiation estimate of the windowed Tx signal (K6).
//
// Replaces wOFDMSystem.estimate_obr's signal construction and __psd_estimate
// (python/ofdm_utils/timefreq_simulation.py:104-123, 216-253): a record of n_sym OFDM symbols on the sub-carrier
// allocation of :223-232 (DC and the 2*guard_band - 1 centre bins carry nothing), IDFT, CP/CS, Tx window, overlap-add
// of the tails (:85-101; tail 0 = plain serialisation), cut into slices of fft_len = 8N samples (the last, partial one
// zero padded), |fftshift(FFT)|^2 averaged over the slices.  The reference evaluates ONE record; here `records`
// independent records are averaged (records = 1 with injected symbols reproduces the reference's estimator).
// One CTA works on two neighbouring slices at a time: the symbols that touch them are drawn (Philox, keyed by record
// and symbol: the result does not depend on the slicing or the grid), transformed with the register IFFT, windowed
// and added into the two slice buffers in shared memory; two 8N-point register FFTs (128 threads each); the squared
// magnitudes stay in registers (each thread owns 16 bins) over all of the CTA's slices and meet in one atomicAdd(double)
// per bin and CTA at the end.  fp32 arithmetic, fp64 accumulation.
#include <algorithm>
#include <cstring>

#include "host_common.h"

namespace wofdm {
namespace {

struct PsdParams {
    int N, cp, cs, tail_tx, bits, n_tx, stride, constellation, guard_band, n_sym;
    long long len;                 // samples of a record's stream: tail_tx + n_sym * stride
    int n_slices;                  // floor(len / P) + 1 (the reference always adds the remainder slice)
    long long records;
    const float* win_tx;           // [n_tx], v_tx * qam_scale / N
    const float2* tw;              // twiddles of the N-point transform
    const float2* twp;             // twiddles of the P-point transform
    unsigned long long seed;
    const int32_t* sym_idx;        // [records][n_sym][N - 2*guard_band] injected constellation indices, or NULL
    double* psd;                   // [P], sum over slices and records
};

// sub-carrier allocation (timefreq_simulation.py:223-232): bin 0 and the bins N/2-gb+1 .. N/2+gb-1 are null; data row
// of an active bin
__device__ __forceinline__ int data_row(int k, int N, int gb) {
    if (k >= 1 && k <= N / 2 - gb) return k - 1;
    if (k >= N / 2 + gb) return k - 2 * gb;
    return -1;
}

template <int N>
__global__ void __launch_bounds__(256, 2) psd_kernel(const PsdParams p) {
    using PN = FftPlan<N>;
    constexpr int P = 8 * N;
    using PP = FftPlan<P>;
    constexpr int NT = 256, TPF = PN::TPF, FPP = NT / TPF, TPFP = PP::TPF, FPPP = NT / TPFP;
    static_assert(FPPP == 2, "two slices per pass");
    constexpr int XCH = (FPP * PN::XLEN > FPPP * PP::XLEN) ? FPP * PN::XLEN : FPPP * PP::XLEN;
    extern __shared__ __align__(16) unsigned char psm[];
    float2* buf = reinterpret_cast<float2*>(psm);          // [2 P] the two slices being built
    float2* xb = buf + 2 * P;                              // FFT exchange (small and big transforms in turn)
    float2* tw = xb + XCH;
    float2* twp = tw + PN::NTW;
    float2* qlut = twp + PP::NTW;
    float* wtx = reinterpret_cast<float*>(qlut + 256);
    const int tid = threadIdx.x, slot = tid / TPF, t = tid % TPF, slotp = tid / TPFP, tp = tid % TPFP;
    const int hb = p.bits >> 1, m = 1 << hb;
    const int n_tx = p.n_tx, stride = p.stride, nd = N - 2 * p.guard_band;
    for (int i = tid; i < PN::NTW; i += NT) tw[i] = p.tw[i];
    for (int i = tid; i < PP::NTW; i += NT) twp[i] = p.twp[i];
    for (int i = tid; i < n_tx; i += NT) wtx[i] = p.win_tx[i];
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
    draw.seed = p.seed; philox_round_keys(p.seed, draw.rk); draw.bits = p.bits;

    double acc[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) acc[q] = 0.0;
    const int pairs = (p.n_slices + 1) / 2;
    const long long items = p.records * pairs;
    for (long long it = blockIdx.x; it < items; it += gridDim.x) {
        const long long rec = it / pairs;
        const int pr = (int)(it - rec * pairs);
        const long long a0 = (long long)pr * 2 * P;        // first stream sample of the pair of slices
        for (int i = tid; i < 2 * P; i += NT) buf[i] = make_float2(0.f, 0.f);
        __syncthreads();
        // symbols that reach [a0, a0 + 2P): s*stride + n_tx > a0 and s*stride < a0 + 2P
        const long long s_lo = a0 >= n_tx ? (a0 - n_tx) / stride + 1 : 0;
        const long long s_end = (a0 + 2 * P - 1) / stride;
        const long long s_hi = s_end < p.n_sym - 1 ? s_end : p.n_sym - 1;
        for (long long s0 = s_lo; s0 <= s_hi; s0 += FPP) {
            const long long s = s0 + slot;
            const bool live = s <= s_hi;
            const long long se = live ? s : s_hi;
            float2 v[16];
            if (p.sym_idx != nullptr) {
                const int32_t* src = p.sym_idx + ((size_t)rec * p.n_sym + (size_t)se) * nd;
#pragma unroll
                for (int q = 0; q < 16; ++q) {
                    const int d = data_row(t + q * TPF, N, p.guard_band);
                    v[q] = d >= 0 ? qlut[src[d] & 0xff] : make_float2(0.f, 0.f);
                }
            } else {
                uint32_t w[4];
                load_sym_idx<N, false>(draw, rec, (int)se, t, w);
#pragma unroll
                for (int q = 0; q < 16; ++q)
                    v[q] = data_row(t + q * TPF, N, p.guard_band) >= 0 ? qlut[sym_byte(w, q)] : make_float2(0.f, 0.f);
            }
            fft_regs<float, N, +1, FPP>(v, t, xb + slot * PN::XLEN, tw, slot);
            // sample i of symbol s = wtx[i] * x[(i - cp) mod N] at stream position s*stride + i; overlapping tails of
            // neighbouring symbols meet in the buffer (two addends per sample at most: order does not matter)
            if (live) {
                const long long base = se * stride - a0;
#pragma unroll
                for (int q = 0; q < 16; ++q) {
                    const int n = t + q * TPF;             // IDFT output index
                    for (int i = n + p.cp; i >= 0; i -= N) {          // i = n + cp (body / suffix wrap handled below) and its prefix copy
                        if (i < n_tx) {
                            const long long pos = base + i;
                            if (pos >= 0 && pos < 2 * P) {
                                const float2 y = cscale(wtx[i], v[q]);
                                atomicAdd(&buf[pos].x, y.x);
                                atomicAdd(&buf[pos].y, y.y);
                            }
                        }
                    }
                    const int is = n + p.cp + N;                       // suffix copy
                    if (is < n_tx) {
                        const long long pos = base + is;
                        if (pos >= 0 && pos < 2 * P) {
                            const float2 y = cscale(wtx[is], v[q]);
                            atomicAdd(&buf[pos].x, y.x);
                            atomicAdd(&buf[pos].y, y.y);
                        }
                    }
                }
            }
            __syncthreads();
        }
        // two P-point transforms; slice 2 pr + slotp exists iff it is < n_slices (its samples past len are zeros)
        {
            float2 v[16];
#pragma unroll
            for (int q = 0; q < 16; ++q) v[q] = buf[slotp * P + tp + q * TPFP];
            fft_regs<float, P, -1, FPPP>(v, tp, xb + slotp * PP::XLEN, twp, slotp);
            if (2 * pr + slotp < p.n_slices) {
#pragma unroll
                for (int q = 0; q < 16; ++q) acc[q] += (double)(v[q].x * v[q].x + v[q].y * v[q].y);
            }
        }
        __syncthreads();
    }
    // fftshift: bin k of the transform is entry (k + P/2) mod P of the estimate
#pragma unroll
    for (int q = 0; q < 16; ++q) atomicAdd(&p.psd[(tp + q * TPFP + P / 2) & (P - 1)], acc[q]);
}

template <int N> size_t psd_smem(int n_tx) {
    using PN = FftPlan<N>;
    using PP = FftPlan<8 * N>;
    constexpr int FPP = 256 / PN::TPF, FPPP = 256 / PP::TPF;
    constexpr int XCH = (FPP * PN::XLEN > FPPP * PP::XLEN) ? FPP * PN::XLEN : FPPP * PP::XLEN;
    return ((size_t)2 * 8 * N + XCH + PN::NTW + PP::NTW + 256) * sizeof(float2) + (((size_t)n_tx + 3) & ~(size_t)3) * sizeof(float);
}

}  // namespace
}  // namespace wofdm

using namespace wofdm;

extern "C" int wofdm_psd_estimate(wofdm_handle h, int N, int cp, int cs, int tail_tx, int bits, int constellation,
                                  const double* win_tx, int guard_band, int n_sym, int64_t records, uint64_t seed,
                                  const int32_t* sym_idx, double* psd) {
    NvtxRange nvtx_("wofdm_psd_estimate");
    if (!h) return WOFDM_EINVAL;
    if (N != 256) return fail(h, WOFDM_EUNSUPPORTED, "the PSD estimate is built for N = 256 (2048-point periodogram, two per CTA)");
    if (!win_tx || !psd) return fail(h, WOFDM_EINVAL, "NULL buffer");
    if (cp < 0 || cp > N || cs < 0 || cs > N || tail_tx < 0 || 2 * tail_tx > N + cp + cs) return fail(h, WOFDM_EINVAL, "bad cp / cs / tail_tx");
    if (bits != 2 && bits != 4 && bits != 6 && bits != 8) return fail(h, WOFDM_EINVAL, "bits must be 2, 4, 6 or 8");
    if (constellation != 0 && constellation != 1) return fail(h, WOFDM_EINVAL, "constellation must be 0 or 1");
    if (guard_band < 1 || 2 * guard_band >= N) return fail(h, WOFDM_EINVAL, "guard_band must lie in [1, N/2)");
    if (n_sym < 1 || records < 1) return fail(h, WOFDM_EINVAL, "n_sym and records must be >= 1");
    const int n_tx = N + cp + cs, stride = n_tx - tail_tx, P = 8 * N;
    const long long len = (long long)tail_tx + (long long)n_sym * stride;
    if (len < P) return fail(h, WOFDM_EINVAL, "record shorter than one periodogram slice (the reference fails there too)");
    DeviceCtx& d = h->devs[0];
    WOFDM_CUDA(h, cudaSetDevice(d.dev));
    wofdm_sys_t sc;
    memset(&sc, 0, sizeof(sc));
    sc.bits = bits; sc.constellation = constellation;
    const double k = qam_scale(sc) / (double)N;            // IDFT 1/N (transmitter.py:58) and constellation scale
    std::vector<float> w(n_tx);
    for (int i = 0; i < n_tx; ++i) w[i] = (float)(win_tx[i] * k);
    std::vector<double> twn = build_twiddles(N), twpd = build_twiddles(P);
    std::vector<float> twf(twn.begin(), twn.end()), twpf(twpd.begin(), twpd.end());
    const size_t n_idx = sym_idx ? (size_t)records * n_sym * (N - 2 * guard_band) : 0;
    int rc = arena_reserve(h, d, w.size() * 4 + twf.size() * 4 + twpf.size() * 4 + n_idx * 4 + (size_t)P * 8 + 256);
    if (rc) return rc;
    float* d_w = static_cast<float*>(arena_take(d, w.size() * 4));
    float* d_tw = static_cast<float*>(arena_take(d, twf.size() * 4));
    float* d_twp = static_cast<float*>(arena_take(d, twpf.size() * 4));
    int32_t* d_idx = n_idx ? static_cast<int32_t*>(arena_take(d, n_idx * 4)) : nullptr;
    double* d_psd = static_cast<double*>(arena_take(d, (size_t)P * 8));
    if (!d_w || !d_tw || !d_twp || !d_psd || (n_idx && !d_idx)) return fail(h, WOFDM_ENOMEM, "arena exhausted");
    WOFDM_CUDA(h, cudaMemcpyAsync(d_w, w.data(), w.size() * 4, cudaMemcpyHostToDevice, d.stream));
    WOFDM_CUDA(h, cudaMemcpyAsync(d_tw, twf.data(), twf.size() * 4, cudaMemcpyHostToDevice, d.stream));
    WOFDM_CUDA(h, cudaMemcpyAsync(d_twp, twpf.data(), twpf.size() * 4, cudaMemcpyHostToDevice, d.stream));
    if (n_idx) WOFDM_CUDA(h, cudaMemcpyAsync(d_idx, sym_idx, n_idx * 4, cudaMemcpyHostToDevice, d.stream));
    WOFDM_CUDA(h, cudaMemsetAsync(d_psd, 0, (size_t)P * 8, d.stream));
    PsdParams p;
    memset(&p, 0, sizeof(p));
    p.N = N; p.cp = cp; p.cs = cs; p.tail_tx = tail_tx; p.bits = bits; p.n_tx = n_tx; p.stride = stride;
    p.constellation = constellation; p.guard_band = guard_band; p.n_sym = n_sym;
    p.len = len; p.n_slices = (int)(len / P) + 1; p.records = records;
    p.win_tx = d_w; p.tw = reinterpret_cast<const float2*>(d_tw); p.twp = reinterpret_cast<const float2*>(d_twp);
    p.seed = seed; p.sym_idx = d_idx; p.psd = d_psd;
    const long long items = records * ((p.n_slices + 1) / 2);
    const int grid = (int)std::min<long long>(items, 2LL * d.sm_count);
    cudaError_t e = cudaSuccess;
    {
        const size_t sm = psd_smem<256>(n_tx);
        e = cudaFuncSetAttribute(psd_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        if (e == cudaSuccess) psd_kernel<256><<<grid, 256, sm, d.stream>>>(p);
    }
    WOFDM_CUDA(h, e);
    WOFDM_CUDA(h, cudaGetLastError());
    h->launches += 1;
    WOFDM_CUDA(h, cudaMemcpyAsync(psd, d_psd, (size_t)P * 8, cudaMemcpyDeviceToHost, d.stream));
    WOFDM_CUDA(h, cudaStreamSynchronize(d.stream));
    const double inv = 1.0 / ((double)p.n_slices * (double)records);
    for (int i = 0; i < P; ++i) psd[i] *= inv;
    return WOFDM_OK;
}
