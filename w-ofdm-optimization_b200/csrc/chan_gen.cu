// chan_gen.cu -- K4, ITU-R tapped-delay-line channels with GMEDS_1 Rayleigh fading on the device.
//
// Replaces channel_model.gen_chan (python/channel_model/itur_channels.py:33-94) and rayleigh_fading_gmeds_1
// (python/channel_model/rayleigh_fading.py:52-102) for the "10k synthetic multipath channels" of BASELINE configs[4]:
// one CTA = one independent SET of channels (one call of the reference's gen_chan: its own random oscillator phases,
// no_frames time samples spaced frame_duration apart); wofdm_optimization.py:77-82 calls it with no_frames = 1 once
// per stored channel, which is n_sets = C, no_frames = 1 here.
//   waveform_p[c] = sqrt(2/K) * sum_k ( cos(2 pi f_re(p,k) t_c + pi a_pk) + i cos(2 pi f_im(p,k) t_c + pi b_pk) ),  K = 21
//   coefficient_p[c] = sqrt(P_p / sum_c |waveform_p[c]|^2) * waveform_p[c]      (energy, not mean power: itur_channels.py:69-74)
//   taps[l][c] = sum_p sinc(tau_p * fs - axis[l]) * coefficient_p[c]
// fp64 throughout, operation order as numpy evaluates it (no FMA contraction in the phase arguments), fixed-order
// reductions: results do not depend on the launch.  Phases: injected (verify) or Philox4x32-10 + Box-Muller (production).
#include <algorithm>
#include <cmath>
#include <cstring>
#include "host_common.h"
#include "common.cuh"

namespace wofdm {

constexpr int CG_OSC = 21;          // oscillators per waveform (itur_channels.py:76)
constexpr int CG_MAXP = 8;          // paths per profile (ITU-R profiles have 4 or 6)
constexpr int CG_THREADS = 128;
enum { STREAM_CHAN = 0x40000000 };

struct ChanGenParams {
    int n_paths, L, no_frames;
    double fd, sampling_freq;        // Doppler frequency; 1 / frame_duration
    double power_lin[CG_MAXP];
    const double* sinc;              // [L][n_paths]
    const double* phases_in;         // [set][path][osc][2] or NULL
    unsigned long long seed;
    double2* out;                    // [set][frame][L]
    // long records (no_frames > CG_FRAMES_PER_BLOCK): a set is spread over n_blk CTAs, energies meet in `partial`
    int n_blk;
    double* partial;                 // [set][n_blk][CG_MAXP]
};
constexpr int CG_FRAMES_PER_BLOCK = 4 * CG_THREADS;

__device__ __forceinline__ double2 chan_waveform(const ChanGenParams& p, const double* ph, int w, double t) {
    const double pi = 3.141592653589793;
    double re = 0.0, im = 0.0;
    for (int o = 0; o < CG_OSC; ++o) {
        const double rot = __dmul_rn(pi / (4 * CG_OSC), (double)w / (double)(p.n_paths + 2));
        const double arr = __dmul_rn(pi / (2 * CG_OSC), (double)o + .5);
        const double f_re = __dmul_rn(p.fd, cos(__dadd_rn(arr, rot))), f_im = __dmul_rn(p.fd, cos(__dadd_rn(arr, -rot)));
        // ((2*pi)*f)*t + pi*phase, rounded after every operation as numpy does
        const double a_re = __dadd_rn(__dmul_rn(__dmul_rn(2 * pi, f_re), t), __dmul_rn(pi, ph[(w * CG_OSC + o) * 2]));
        const double a_im = __dadd_rn(__dmul_rn(__dmul_rn(2 * pi, f_im), t), __dmul_rn(pi, ph[(w * CG_OSC + o) * 2 + 1]));
        re = __dadd_rn(re, cos(a_re));
        im = __dadd_rn(im, cos(a_im));
    }
    const double sc = sqrt(2.0 / CG_OSC);
    return make_double2(sc * re, sc * im);
}

// PHASE 0: one CTA per set does everything (the reference's use: short sets).  Long records run twice over a grid of
// (set, frame block): PHASE 1 leaves every block's path energies in p.partial, PHASE 2 adds them in block order
// (every CTA of the set forms the same scale) and writes the block's taps.
template <int PHASE>
__global__ void __launch_bounds__(CG_THREADS) chan_gen_kernel(const ChanGenParams p) {
    extern __shared__ __align__(16) unsigned char sm_raw[];
    double* ph = reinterpret_cast<double*>(sm_raw);                       // [n_paths][21][2]
    double* red = ph + CG_MAXP * CG_OSC * 2;                              // [CG_THREADS]
    double* scale = red + CG_THREADS;                                     // [n_paths]
    const int tid = threadIdx.x;
    const long long set = blockIdx.x;
    const int c_lo = PHASE == 0 ? 0 : blockIdx.y * CG_FRAMES_PER_BLOCK;
    const int c_hi = PHASE == 0 ? p.no_frames : min(p.no_frames, c_lo + CG_FRAMES_PER_BLOCK);
    const int nph = p.n_paths * CG_OSC;
    for (int q = tid; q < nph; q += CG_THREADS) {
        if (p.phases_in) {
            ph[2 * q] = p.phases_in[((size_t)set * nph + q) * 2];
            ph[2 * q + 1] = p.phases_in[((size_t)set * nph + q) * 2 + 1];
        } else {
            const uint4 r = philox4x32_10(make_uint4((uint32_t)set, (uint32_t)((unsigned long long)set >> 32), (uint32_t)q, STREAM_CHAN),
                                          (uint32_t)p.seed, (uint32_t)(p.seed >> 32));
            const double2 g = gauss_pair(r.x, r.y, double());
            ph[2 * q] = g.x;
            ph[2 * q + 1] = g.y;
        }
    }
    __syncthreads();
    // energy of every path's waveform over the set's frames: per-thread partials in frame order, fixed-order tree
    if (PHASE != 2) {
        for (int w = 0; w < p.n_paths; ++w) {
            double e = 0.0;
            for (int c = c_lo + tid; c < c_hi; c += CG_THREADS) {
                const double2 v = chan_waveform(p, ph, w, (double)c / p.sampling_freq);
                e += v.x * v.x + v.y * v.y;
            }
            red[tid] = e;
            __syncthreads();
            for (int s = CG_THREADS / 2; s > 0; s >>= 1) {
                if (tid < s) red[tid] += red[tid + s];
                __syncthreads();
            }
            if (tid == 0) {
                if (PHASE == 0) scale[w] = sqrt(p.power_lin[w] / red[0]);
                else p.partial[((size_t)set * p.n_blk + blockIdx.y) * CG_MAXP + w] = red[0];
            }
            __syncthreads();
        }
        if (PHASE == 1) return;
    } else {
        if (tid < p.n_paths) {
            double e = 0.0;
            for (int b = 0; b < p.n_blk; ++b) e += p.partial[((size_t)set * p.n_blk + b) * CG_MAXP + tid];
            scale[tid] = sqrt(p.power_lin[tid] / e);
        }
        __syncthreads();
    }
    for (int c = c_lo + tid; c < c_hi; c += CG_THREADS) {
        double2 coef[CG_MAXP];
        for (int w = 0; w < p.n_paths; ++w) {
            const double2 v = chan_waveform(p, ph, w, (double)c / p.sampling_freq);
            coef[w] = make_double2(scale[w] * v.x, scale[w] * v.y);
        }
        double2* dst = p.out + ((size_t)set * p.no_frames + c) * p.L;
        for (int l = 0; l < p.L; ++l) {
            double re = 0.0, im = 0.0;
            for (int w = 0; w < p.n_paths; ++w) { const double s = p.sinc[l * p.n_paths + w]; re += s * coef[w].x; im += s * coef[w].y; }
            dst[l] = make_double2(re, im);
        }
    }
}

namespace {
struct Profile { const char* name; int n; double delay[CG_MAXP]; double power_db[CG_MAXP]; };
// python/channel_model/itur_channels.py:14-30
const Profile PROFILES[4] = {
    {"vehicularA", 6, {0, 310e-9, 710e-9, 1090e-9, 1730e-9, 2510e-9}, {0, -1, -9, -10, -15, -20}},
    {"vehicularB", 6, {0, 300e-9, 8900e-9, 12900e-9, 17100e-9, 20000e-9}, {-2.5, 0, -12.8, -10, -25.2, -16}},
    {"outdoor-indoorA", 4, {0, 110e-9, 190e-9, 410e-9}, {0, -9.7, -19.2, -22.8}},
    {"outdoor-indoorB", 6, {0, 200e-9, 800e-9, 1200e-9, 2300e-9, 3700e-9}, {0, -.9, -4.9, -8, -7.8, -23.9}},
};
}  // namespace
}  // namespace wofdm

using namespace wofdm;

extern "C" {

int wofdm_channel_profile(const char* standard) {
    if (!standard) return WOFDM_EINVAL;
    for (int i = 0; i < 4; ++i)
        if (!strcmp(standard, PROFILES[i].name)) return i;
    return WOFDM_EINVAL;
}

int wofdm_gen_channels(wofdm_handle h, int profile, int L, double doppler_freq, double sampling_rate, double frame_duration,
                       int no_frames, int n_sets, uint64_t seed, const double* phases, double* chan) {
    NvtxRange nvtx_("wofdm_gen_channels");
    if (!h) return WOFDM_EINVAL;
    if (profile < 0 || profile > 3) return fail(h, WOFDM_EINVAL, "unknown ITU-R channel profile");
    if (L < 1 || no_frames < 1 || n_sets < 1 || !chan) return fail(h, WOFDM_EINVAL, "bad channel-generation arguments");
    if (!(frame_duration > 0) || !(sampling_rate > 0)) return fail(h, WOFDM_EINVAL, "rates must be positive");
    const Profile& pf = PROFILES[profile];
    DeviceCtx& d = h->devs[0];
    WOFDM_CUDA(h, cudaSetDevice(d.dev));
    // sinc_mat[l][p] = sinc(tau_p * fs - axis[l]), axis = linspace(-(L-1)/2, (L+1)/2, L)   (itur_channels.py:79-83)
    std::vector<double> sinc((size_t)L * pf.n);
    const double start = -(L - 1) / 2.0, stop = (L + 1) / 2.0, step = L > 1 ? (stop - start) / (L - 1) : 0.0;
    for (int l = 0; l < L; ++l) {
        const double ax = (l == L - 1 && L > 1) ? stop : start + l * step;
        for (int p = 0; p < pf.n; ++p) {
            const double x = pf.delay[p] * sampling_rate - ax;
            const double y = 3.141592653589793 * (x == 0.0 ? 1e-20 : x);
            sinc[(size_t)l * pf.n + p] = std::sin(y) / y;
        }
    }
    const size_t n_ph = phases ? (size_t)n_sets * pf.n * CG_OSC * 2 : 0;
    const size_t n_out = (size_t)n_sets * no_frames * L;
    const int n_blk = no_frames > CG_FRAMES_PER_BLOCK ? (no_frames + CG_FRAMES_PER_BLOCK - 1) / CG_FRAMES_PER_BLOCK : 1;
    const size_t n_part = n_blk > 1 ? (size_t)n_sets * n_blk * CG_MAXP : 0;
    int rc = arena_reserve(h, d, sinc.size() * 8 + n_ph * 8 + n_out * 16 + n_part * 8);
    if (rc) return rc;
    double* d_sinc = static_cast<double*>(arena_take(d, sinc.size() * 8));
    double* d_ph = n_ph ? static_cast<double*>(arena_take(d, n_ph * 8)) : nullptr;
    double2* d_out = static_cast<double2*>(arena_take(d, n_out * 16));
    double* d_part = n_part ? static_cast<double*>(arena_take(d, n_part * 8)) : nullptr;
    if (!d_sinc || !d_out || (n_ph && !d_ph) || (n_part && !d_part)) return fail(h, WOFDM_ENOMEM, "arena exhausted");
    WOFDM_CUDA(h, cudaMemcpyAsync(d_sinc, sinc.data(), sinc.size() * 8, cudaMemcpyHostToDevice, d.stream));
    if (n_ph) WOFDM_CUDA(h, cudaMemcpyAsync(d_ph, phases, n_ph * 8, cudaMemcpyHostToDevice, d.stream));
    ChanGenParams p;
    memset(&p, 0, sizeof(p));
    p.n_paths = pf.n; p.L = L; p.no_frames = no_frames;
    p.fd = doppler_freq; p.sampling_freq = 1.0 / frame_duration;
    for (int i = 0; i < pf.n; ++i) p.power_lin[i] = std::pow(10.0, pf.power_db[i] / 10.0);
    p.sinc = d_sinc; p.phases_in = d_ph; p.seed = seed; p.out = d_out;
    const size_t smem = (size_t)(CG_MAXP * CG_OSC * 2 + CG_THREADS + CG_MAXP) * sizeof(double);
    p.n_blk = n_blk; p.partial = d_part;
    if (n_blk == 1) {
        chan_gen_kernel<0><<<n_sets, CG_THREADS, smem, d.stream>>>(p);
        h->launches += 1;
    } else {
        chan_gen_kernel<1><<<dim3(n_sets, n_blk), CG_THREADS, smem, d.stream>>>(p);
        chan_gen_kernel<2><<<dim3(n_sets, n_blk), CG_THREADS, smem, d.stream>>>(p);
        h->launches += 2;
    }
    WOFDM_CUDA(h, cudaGetLastError());
    // [set][frame][L] complex == L x (n_sets*no_frames) column-major
    WOFDM_CUDA(h, cudaMemcpyAsync(chan, d_out, n_out * 16, cudaMemcpyDeviceToHost, d.stream));
    WOFDM_CUDA(h, cudaStreamSynchronize(d.stream));
    return WOFDM_OK;
}

}  // extern "C"
