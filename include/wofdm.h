/* wofdm.h -- C-ABI of libwofdm.so, the B200-native replacement for the Monte-Carlo BER/SER
 * chain and the interference-power evaluation of felipescoelho/w-ofdm-optimization.
 *
 * The reference has no FFI: its boundary is function level.  Each entry point below names the
 * reference function it replaces (paths relative to the reference root).  Bindings: ctypes
 * (w-ofdm-optimization_b200/capi.py) for python/ofdm_utils, MEX gateway (mex/wofdm_mex.cpp) for
 * matlab/main_BER_calculation.m and matlab/main_interference_calculation.m; see INTEGRATION.md.
 *
 * Conventions: plain C, host pointers unless a parameter is named d_*; the caller owns every
 * buffer; complex = interleaved (re,im) doubles; matrices are column-major where MATLAB passes
 * matrices; every function returns 0 or a negative WOFDM_E* code and never throws; a handle is
 * not thread-safe, different handles are.  There is no CPU fallback: without a CUDA device
 * wofdm_create fails with WOFDM_ENODEV.
 */
#ifndef WOFDM_H_
#define WOFDM_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WOFDM_VERSION 100

#if defined(__GNUC__)
#define WOFDM_API __attribute__((visibility("default")))
#else
#define WOFDM_API
#endif

enum {
    WOFDM_OK = 0,
    WOFDM_EINVAL = -1,       /* bad argument (see wofdm_last_error) */
    WOFDM_ECUDA = -2,        /* CUDA runtime error */
    WOFDM_ENOMEM = -3,       /* host or device allocation failed */
    WOFDM_ENODEV = -4,       /* no usable CUDA device */
    WOFDM_EUNSUPPORTED = -5  /* valid request outside the compiled kernel set */
};

typedef struct wofdm_ctx* wofdm_handle;
typedef struct wofdm_ber_plan_s* wofdm_ber_plan;

/* One w-OFDM system.  n_tx = N+cp+cs, stride = n_rx = N+tail_rx+rm = n_tx-tail_tx. */
typedef struct {
    int32_t N;             /* DFT length, power of two in [16, 1024] */
    int32_t cp;            /* cyclic prefix */
    int32_t cs;            /* cyclic suffix */
    int32_t tail_tx;       /* Tx window tail */
    int32_t tail_rx;       /* Rx window tail (even) */
    int32_t rm;            /* samples dropped in front of each Rx block */
    int32_t shift;         /* circular shift before the DFT */
    int32_t bits;          /* bits per sub-carrier: 2, 4, 6 or 8 (square QAM) */
    int32_t S;             /* OFDM symbols per frame; symbol 0 is the pilot */
    int32_t noise_norm;    /* 0 = python: SNR fixed on the truncated signal (wofdm_simulation.py:208-215)
                              1 = matlab: on the full convolution (main_BER_calculation.m:260-261) */
    int32_t constellation; /* 0 = python: natural-order un-normalised list (wofdm_simulation.py:179-182)
                              1 = matlab: qammod Gray, unit average power (main_BER_calculation.m:248) */
    int32_t precision;     /* 0 = fp32, 1 = fp64 */
    int32_t guard;         /* null sub-carriers on EACH side of the centred spectrum (0 = all N carry data):
                              matlab/main_channel_mask.m:55,388-391 (`offset`, zeros + ifftshift).  FFT bin k is
                              active iff guard <= (k + N/2) mod N < N - guard; only active bins are counted.
                              Needs bits < 8. */
} wofdm_sys_t;

/* ---- library / device ------------------------------------------------------------------- */
WOFDM_API int wofdm_version(void);
WOFDM_API int wofdm_device_count(int* n);
/* n_gpus = 0: every visible device; n > 0: devices 0..n-1. */
WOFDM_API int wofdm_create(wofdm_handle* h, int n_gpus);
/* Explicit device list (one process per GPU under torchrun: {LOCAL_RANK}). */
WOFDM_API int wofdm_create_on(wofdm_handle* h, const int* device_ids, int n);
WOFDM_API int wofdm_destroy(wofdm_handle h);
WOFDM_API const char* wofdm_last_error(wofdm_handle h);
/* Kernels of this library launched through the handle so far. */
WOFDM_API int64_t wofdm_launch_count(wofdm_handle h);

/* Diagnostic: FMA-only micro-benchmark on device slot 0 (mode 0 scalar FFMA, 1 packed FFMA2).
 * tflops = measured FP32 rate, the denominator of the BER kernel's roofline; sm_mhz_equiv (optional)
 * = the SM clock at which 128 FMA lanes per SM would deliver that rate. */
WOFDM_API int wofdm_diag_fp32_peak(wofdm_handle h, int mode, double* tflops, double* sm_mhz_equiv);

/* ---- parameter table and windows (host, tiny) -------------------------------------------- */
/* Replaces wOFDMSystem.__init__'s table (python/ofdm_utils/wofdm_simulation.py:391-418) and
 * calculate_parameters (matlab/main_BER_calculation.m:457-493).  name: CP, wtx, CPwtx, wrx, CPwrx,
 * WOLA, CPW.  Fills N, cp, cs, tail_tx, tail_rx, rm, shift; leaves the other fields alone. */
WOFDM_API int wofdm_params_from_name(const char* name, int N, int cp, int tail_tx, int tail_rx, wofdm_sys_t* out);
/* gen_rc_window_tx / gen_rc_window_rx (python/ofdm_utils/transmitter.py:61-87, receiver.py:36-56;
 * matlab/functions/transmitter_rc_window.m, receiver_rc_window.m).  out: n_tx resp. N+tail_rx. */
WOFDM_API int wofdm_rc_window_tx(const wofdm_sys_t* sys, double* out);
WOFDM_API int wofdm_rc_window_rx(const wofdm_sys_t* sys, double* out);
/* reduce_variable_tx/rx applied to a stored tail vector (python/optimization_tools/utils.py:13-73):
 * x has tail_tx+1 resp. tail_rx/2+1 entries. */
WOFDM_API int wofdm_expand_window_tx(const wofdm_sys_t* sys, const double* x, double* out);
WOFDM_API int wofdm_expand_window_rx(const wofdm_sys_t* sys, const double* x, double* out);

/* ---- Monte-Carlo BER/SER chain ----------------------------------------------------------- */
/* Replaces wOFDMSystem.__run_sim_mc / __run_sim_cp_mc (python/ofdm_utils/wofdm_simulation.py:85-366)
 * and run_simulation (matlab/main_BER_calculation.m:230-274) for ONE window pair.  On-device Philox
 * draws.  Frames are indexed f = (snr_idx*C + chan_idx)*ensemble + e; symbols depend on (seed, f),
 * noise on (seed, f, variant): calling twice with the same seed and variant 0/1 evaluates two window
 * pairs on the same symbols with independent noise, as the reference does for optimised vs RC.
 * Counters are summed over channels x ensemble, one entry per SNR point; *_tot are the totals
 * (frames * N*bits*(S-1) resp. frames * N*(S-1)).  Uses every device of the handle.
 *   win_tx: n_tx doubles, win_rx: N+tail_rx doubles, chan: complex L x C column-major. */
WOFDM_API int wofdm_ber_run(wofdm_handle h, const wofdm_sys_t* sys, const double* win_tx, const double* win_rx,
                  const double* chan, int L, int C, const double* snr_db, int n_snr, int64_t ensemble,
                  uint64_t seed, uint32_t variant,
                  int64_t* bit_err, int64_t* bit_tot, int64_t* sym_err, int64_t* sym_tot);
/* Same, but only frames f = shard_index (mod shard_count): one process per GPU sums the results
 * with a single all-reduce of the int64 counters. */
WOFDM_API int wofdm_ber_run_shard(wofdm_handle h, const wofdm_sys_t* sys, const double* win_tx, const double* win_rx,
                        const double* chan, int L, int C, const double* snr_db, int n_snr, int64_t ensemble,
                        uint64_t seed, uint32_t variant, int shard_index, int shard_count,
                        int64_t* bit_err, int64_t* bit_tot, int64_t* sym_err, int64_t* sym_tot);

/* Several window pairs on the SAME symbols in one call -- what the reference's loops do per frame: optimised and RC
 * windows on one signal_digmod with independent noise (python/ofdm_utils/wofdm_simulation.py:183-236), the RC window
 * and the six CaseA/CaseB steps of WOLA / CPW (matlab/main_BER_calculation.m:118-198).
 *   win_tx: n_var x n_tx doubles (pair v at win_tx + v*n_tx), win_rx: n_var x (N+tail_rx), n_var <= WOFDM_MAX_VARIANTS.
 *   bit_err / sym_err: n_var x n_snr (pair v at + v*n_snr); bit_tot / sym_tot: n_snr (the same for every pair).
 * Pair v gets exactly the counters of wofdm_ber_run_shard(..., variant + v, ...) with its windows: symbols depend on
 * (seed, frame), pair v's noise on (seed, frame, variant + v).  Where the tensor-core kernel applies, all pairs of a frame
 * are evaluated by ONE launch (symbols drawn once, tables of all pairs resident); otherwise one launch per pair. */
#define WOFDM_MAX_VARIANTS 8
WOFDM_API int wofdm_ber_run_multi(wofdm_handle h, const wofdm_sys_t* sys, const double* win_tx, const double* win_rx, int n_var,
                        const double* chan, int L, int C, const double* snr_db, int n_snr, int64_t ensemble,
                        uint64_t seed, uint32_t variant, int shard_index, int shard_count,
                        int64_t* bit_err, int64_t* bit_tot, int64_t* sym_err, int64_t* sym_tot);

/* Device-resident form: inputs are uploaded once, launches are asynchronous. */
WOFDM_API int wofdm_ber_plan_create(wofdm_handle h, const wofdm_sys_t* sys, const double* win_tx, const double* win_rx,
                          const double* chan, int L, int C, const double* snr_db, int n_snr,
                          wofdm_ber_plan* plan);
/* The same for n_var window pairs (layout as in wofdm_ber_run_multi): counters are int64[n_var][n_snr][2] and
 * wofdm_ber_plan_read fills n_var x n_snr entries.  wofdm_ber_plan_variants returns n_var; *fused (optional) = 1 if one
 * launch evaluates all pairs. */
WOFDM_API int wofdm_ber_plan_create_multi(wofdm_handle h, const wofdm_sys_t* sys, const double* win_tx, const double* win_rx,
                                int n_var, const double* chan, int L, int C, const double* snr_db, int n_snr,
                                wofdm_ber_plan* plan);
WOFDM_API int wofdm_ber_plan_variants(wofdm_ber_plan plan, int* fused);
/* Zeroes the plan's device counters and launches the shard on device slot `slot` of the handle,
 * on `stream` (a cudaStream_t; NULL = the plan's own stream).  Returns without synchronising.
 * One launch per slot and read: a second launch on the same slot before wofdm_ber_plan_read REPLACES the
 * first launch's counts (it waits for the first one if that ran on a different stream).
 * d_counters (optional out) = device pointer to int64[n_snr][2] = {bit_err, sym_err}. */
WOFDM_API int wofdm_ber_plan_launch(wofdm_ber_plan plan, int slot, int64_t ensemble, uint64_t seed, uint32_t variant,
                          int shard_index, int shard_count, void* stream, void** d_counters);
/* Waits for the launches and adds up the counters of the slots (devices) launched since the last read:
 * each slot contributes its most recent launch. */
WOFDM_API int wofdm_ber_plan_read(wofdm_ber_plan plan, int64_t* bit_err, int64_t* sym_err);
/* Name of the kernel variant the plan dispatches to (for logs / profiles). */
WOFDM_API const char* wofdm_ber_plan_kernel(wofdm_ber_plan plan);
WOFDM_API int wofdm_ber_plan_destroy(wofdm_ber_plan plan);

/* Verify mode: the caller injects every draw for F frames and gets the intermediates back.
 * Same chain and same kernels as wofdm_ber_run.
 *   chan: complex L x F (one channel per frame), snr_db: F, sym_idx: N x S x F constellation indices
 *   in the convention's own order, noise: complex len x F unit-variance draws with
 *   len = S*stride (noise_norm 0) or tail_tx + S*stride + L - 1 (noise_norm 1),
 *   eq_out: complex N x (S-1) x F, dec_idx: N x (S-1) x F, bit_err/sym_err: F.
 *   variant_kernel: 0 = the kernel wofdm_ber_run would pick, 1 = force the generic staged kernel,
 *   2 = the register-resident direct-form convolution over the whole frame (no tensor-core convolution, no
 *   circular interior), 3 = exclude the tensor-core convolution kernels only. */
WOFDM_API int wofdm_ber_verify(wofdm_handle h, const wofdm_sys_t* sys, const double* win_tx, const double* win_rx,
                     const double* chan, int L, int F, const double* snr_db,
                     const int32_t* sym_idx, const double* noise, int variant_kernel,
                     double* eq_out, int32_t* dec_idx, int64_t* bit_err, int64_t* sym_err);
/* Exports the on-device draws of production frames frame_ids[0..F) so a production run can be
 * replayed through wofdm_ber_verify or a CPU oracle.  sym_idx: N x S x F, noise: complex len x F. */
WOFDM_API int wofdm_ber_draws(wofdm_handle h, const wofdm_sys_t* sys, int L, uint64_t seed, uint32_t variant,
                    const int64_t* frame_ids, int F, int32_t* sym_idx, double* noise);

/* ---- interference power ------------------------------------------------------------------ */
/* Replaces interf_power (python/ofdm_utils/interf_calc.py:20-113) per channel realisation:
 * P[k + N*c] = sum_{j!=k} |A_0[k,j]|^2 + sum_{m>=1} sum_j |A_m[k,j]|^2,  A_m = Rx_mat . H_m(h_c) . Tx_mat.
 * mode 0 = fp64 (DMMA), one contraction per channel and slice; 1 = TF32-split tensor cores (fp32 grade); 2 = fp64,
 * Hermitian form in the taps: A_m(c) = sum_l h_c[l] G_{m,l}, so the L impulse responses G go through the mode-0
 * contraction once per window pair and every channel costs L^2 MACs per sub-carrier (P_k = h^H Q_k h) -- the mode for
 * many channels (C > L; L <= 88).  chan: complex L x C, P: N x C column-major. */
WOFDM_API int wofdm_interf_power(wofdm_handle h, const wofdm_sys_t* sys, const double* win_tx, const double* win_rx,
                       const double* chan, int L, int C, int mode, double* P);
/* MATLAB semantics (calculate_interference, matlab/main_interference_calculation.m:177-225):
 * scalar per channel, ISI slices summed before squaring.  P: C doubles. */
WOFDM_API int wofdm_interf_power_scalar(wofdm_handle h, const wofdm_sys_t* sys, const double* win_tx,
                              const double* win_rx, const double* chan, int L, int C, int mode, double* P);

/* Device times of the handle's last wofdm_interf_power* call, from CUDA events on its stream (ms; -1 = not taken:
 * kernel times exist for calls whose channels fit one batch): the whole call on the device (uploads, matrix builders,
 * band product, contraction, download), the band product B = H.Tx_mat alone, the tensor-core contraction alone.
 * k_slice0 / k_isi: rows of the real-form K dimension contracted for slice 0 / for every ISI slice (only the non-zero
 * prefix of an ISI slice exists), i.e. the executed GEMM is 2 * 2N * k * N flop per (channel, slice).  Any may be NULL. */
WOFDM_API int wofdm_interf_last_timing(wofdm_handle h, double* total_ms, double* band_ms, double* gemm_ms, int* k_slice0, int* k_isi);

/* ---- Channel-mask BER variant (next-row 8f-1) ----------------------------------------------------------------
 * run_sim_mc of matlab/main_channel_mask.m:334-360 for the MASKED signal: the guard-band frame (sys->guard = offset)
 * whose windowed symbols pass the DFT-domain raised-cosine mask of length 2 n_tx - 1 (dft_rc_filt, :398-417, roll_off
 * as `rollOff`, 10 in the reference); the filter tail of each symbol is added to the next one.  Same arguments, draws
 * (the symbol stream does not depend on `variant`) and counters as wofdm_ber_run, whose result with the same sys is
 * the UNMASKED ber of that script.  fp32, N = 128 / 256 / 512; the frames are split over the devices of the handle
 * (contiguous ranges of the global frame id: counters do not depend on the split).  The Tx side is one dense tensor-core
 * product (csrc/mask_gemm.cu: the script's matrices multiplied out once); WOFDM_MASK_FFT=1 in the environment selects
 * the per-symbol FFT kernel (csrc/mask_kernel.cuh) instead. */
WOFDM_API int wofdm_ber_run_masked(wofdm_handle h, const wofdm_sys_t* sys, const double* win_tx, const double* win_rx,
                         const double* chan, int L, int C, const double* snr_db, int n_snr, int64_t ensemble,
                         uint64_t seed, uint32_t variant, int roll_off,
                         int64_t* bit_err, int64_t* bit_tot, int64_t* sym_err, int64_t* sym_tot);

/* ---- Window-optimisation Hessian (next-row 8f-2) -------------------------------------------------------------
 * The quadratic form of the interference power in the reduced window variables: OptimizerTx / OptimizerRx /
 * OptimizerTxRx.gen_hessian (python/optimization_tools/optimizers.py:232-257, 387-410, 808-833) and
 * quad_objective_tx / _rx (matlab/window_optimization.m:596-680), whose O(n^2 N^2) loops are the slowest part of the
 * reference's run_opt.  chan: ONE impulse response, complex L (the reference passes the mean of the stored set).
 * H: n_var x n_var doubles (symmetric), n_var = (tail_rx/2 + 1) * (tail_tx + 1), variable index
 * u = a*(tail_tx+1) + b for Rx-tail variable a and Tx-tail variable b (OptimizerTxRx's flatten order; Tx-only and
 * Rx-only systems have tail_rx = 0 or tail_tx = 0 and the index is b or a).  *n_var (may be NULL) returns n_var. */
WOFDM_API int wofdm_window_hessian(wofdm_handle h, const wofdm_sys_t* sys, const double* chan, int L, double* H, int* n_var);
/* The two parts of that quadratic form for ARBITRARY basis windows: basis_tx = n_tb windows of n_tx samples, basis_rx = n_rb
 * windows of N + tail_rx samples (row-major), variable u = a*n_tb + b, n_var = n_rb*n_tb:
 *   H_ici[u,u'] = 2 Re <offdiag A0_u, offdiag A0_u'>,   H_isi[u,u'] = 2 Re <AS_u, AS_u'>   (wofdm_window_hessian = their sum
 * on the reduced bases).  With the FULL window as the variable (basis = identity) and the other side's window fixed (one
 * row) these are quad_objective_tx / _rx of matlab/window_optimization.m:596-680:
 *   HTx (or HRx) = alpha * H_ici + (1 - alpha) * diag(row sums of H_isi)       (H2 = real(diag(diag(B2' B2 C C'))), :628-631). */
WOFDM_API int wofdm_window_hessian_parts(wofdm_handle h, const wofdm_sys_t* sys, const double* chan, int L,
                               const double* basis_tx, int n_tb, const double* basis_rx, int n_rb, double* H_ici, double* H_isi);

/* ---- Channel generation (next-row 8f-4) ------------------------------------------------------------------
 * ITU-R tapped-delay-line channels with GMEDS_1 Rayleigh fading on the device: channel_model.gen_chan
 * (python/channel_model/itur_channels.py:33-94 + rayleigh_fading.py:52-102), the producer of channels/<standard>.npy
 * (python/wofdm_optimization.py:63-86).
 *   wofdm_channel_profile: "vehicularA" | "vehicularB" | "outdoor-indoorA" | "outdoor-indoorB" -> index, or WOFDM_EINVAL.
 *   One SET = one call of the reference's gen_chan: its own oscillator phases, no_frames time samples frame_duration
 *   apart, every path scaled to ENERGY 10^(dB/10) over the set (the reference's adjust_power).  The reference's
 *   driver stores C channels as C sets of one frame each: n_sets = C, no_frames = 1.
 *   phases: NULL -> Philox4x32-10 + Box-Muller draws keyed by (seed, set); else the standard-normal draws themselves,
 *   [n_sets][n_paths][21][2] in the reference's draw order (path, oscillator, real part / imaginary part).
 *   chan: complex L x (n_sets*no_frames) column-major, column = set*no_frames + frame. */
WOFDM_API int wofdm_channel_profile(const char* standard);
WOFDM_API int wofdm_gen_channels(wofdm_handle h, int profile, int L, double doppler_freq, double sampling_rate,
                       double frame_duration, int no_frames, int n_sets, uint64_t seed, const double* phases,
                       double* chan);

/* Next row 8f-3: PSD / out-of-band-radiation estimate of the windowed Tx signal.  Replaces the signal construction of
 * wOFDMSystem.estimate_obr and __psd_estimate (python/ofdm_utils/timefreq_simulation.py:104-123, 216-253; MATLAB twin
 * matlab/main_OOB_figures.m:121-159): records of n_sym OFDM symbols on the sub-carrier allocation of :223-232 (DC and the
 * 2*guard_band - 1 centre bins null, N - 2*guard_band data rows), IDFT, CP/CS, Tx window win_tx (N + cp + cs values),
 * overlap-add of the tails (tail_tx = 0: plain serialisation, as the reference does for the un-windowed signals), cut
 * into slices of 8N samples (the last, partial one zero padded), |fftshift(FFT)|^2 averaged over slices and records.
 *   sym_idx: int32 [records][n_sym][N - 2*guard_band] injected constellation indices (records = 1 reproduces the
 *   reference's estimator for its draw), or NULL: on-device Philox draws keyed by (seed, record, symbol).
 *   psd: 8N doubles, the reference's X_est (fftshifted).  N = 256. */
WOFDM_API int wofdm_psd_estimate(wofdm_handle h, int N, int cp, int cs, int tail_tx, int bits, int constellation,
                       const double* win_tx, int guard_band, int n_sym, int64_t records, uint64_t seed,
                       const int32_t* sym_idx, double* psd);

#ifdef __cplusplus
}
#endif
#endif /* WOFDM_H_ */
