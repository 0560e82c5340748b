// wofdm_mex.cpp -- MATLAB MEX gateway over libwofdm.so (include/wofdm.h).  Pure argument marshalling.
//
//   ber = wofdm_mex('run_simulation', ensemble, symbolsPerTx, bitsPerSubcarrier, numSubcar, cpLength, ...
//                   csLength, windowTx, channel, snr, tailTx, tailRx, windowRx, prefixRemovalLength, ...
//                   circularShiftLength [, seed])
//       -- the positional arguments of run_simulation, matlab/main_BER_calculation.m:230-232.
//   berSNR = wofdm_mex('run_simulation_sweep', <the same 14 arguments> [, seed])
//       -- channel is now the whole C x L matrix (rows = realisations) and snr a vector: the triple loop of the
//          driver, matlab/main_BER_calculation.m:66-87 (for snr, for channel: ber = ber + run_simulation(...);
//          berSNR(snrIndex) = ber/numChannels), as ONE device job.  Returns n_snr x 1.
//   P   = wofdm_mex('calculate_interference', cpLength, typeOFDM, windowTx, windowRx, numSubcar, tailTx, ...
//                   tailRx, channels [, mode])
//       -- calculate_interference, matlab/main_interference_calculation.m:177-180; the values the reference
//          reads from settingsData.mat / the channel file inside the function are passed explicitly.
// Windows arrive as dense diagonal matrices (or vectors); channels as rows = realisations.
#ifdef WOFDM_MEX_STUB
#include "mex_stub.h"
#else
#include "mex.h"
#endif
#include <string.h>

#include <vector>

#include "../include/wofdm.h"

static wofdm_handle g_handle = nullptr;

static void at_exit() {
    if (g_handle) wofdm_destroy(g_handle);
    g_handle = nullptr;
}

static wofdm_handle handle() {
    if (!g_handle) {
        if (wofdm_create(&g_handle, 0) != WOFDM_OK)
            mexErrMsgIdAndTxt("wofdm:nodevice", "wofdm_create failed: no CUDA device (this library has no CPU path)");
        mexAtExit(at_exit);
    }
    return g_handle;
}

static void check(int rc, const char* what) {
    if (rc != WOFDM_OK) mexErrMsgIdAndTxt("wofdm:error", "%s failed (%d): %s", what, rc, wofdm_last_error(g_handle));
}

// diagonal of a square matrix, or the vector itself
static std::vector<double> diag_of(const mxArray* a, size_t want) {
    const size_t m = mxGetM(a), n = mxGetN(a);
    const double* p = mxGetPr(a);
    std::vector<double> d(want);
    if (m == want && n == want) {
        for (size_t i = 0; i < want; ++i) d[i] = p[i * (want + 1)];
    } else if (m * n == want) {
        for (size_t i = 0; i < want; ++i) d[i] = p[i];
    } else {
        mexErrMsgIdAndTxt("wofdm:size", "window must be a %d x %d diagonal matrix or a vector of that length", (int)want, (int)want);
    }
    return d;
}

// MATLAB complex (split storage) -> interleaved (re, im)
static std::vector<double> interleave(const mxArray* a, size_t count, size_t stride, size_t offset) {
    const double *re = mxGetPr(a), *im = mxIsComplex(a) ? mxGetPi(a) : nullptr;
    std::vector<double> z(2 * count);
    for (size_t i = 0; i < count; ++i) {
        z[2 * i] = re[offset + i * stride];
        z[2 * i + 1] = im ? im[offset + i * stride] : 0.0;
    }
    return z;
}

// sweep = false: one channel (vector), one SNR -> scalar; sweep = true: C x L channel matrix, SNR vector -> n_snr x 1
static void do_run_simulation(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[], bool sweep) {
    if (nrhs < 14) mexErrMsgIdAndTxt("wofdm:nargin", "run_simulation needs 14 arguments");
    (void)nlhs;
    wofdm_sys_t s;
    memset(&s, 0, sizeof(s));
    const long long ensemble = (long long)mxGetScalar(prhs[0]);
    s.S = (int)mxGetScalar(prhs[1]);
    s.bits = (int)mxGetScalar(prhs[2]);
    s.N = (int)mxGetScalar(prhs[3]);
    s.cp = (int)mxGetScalar(prhs[4]);
    s.cs = (int)mxGetScalar(prhs[5]);
    s.tail_tx = (int)mxGetScalar(prhs[9]);
    s.tail_rx = (int)mxGetScalar(prhs[10]);
    s.rm = (int)mxGetScalar(prhs[12]);
    s.shift = (int)mxGetScalar(prhs[13]);
    s.noise_norm = 1;      // add_wgn on the full convolution, main_BER_calculation.m:260,277-294
    s.constellation = 1;   // qammod(..., 'UnitAveragePower', true), :248
    s.precision = 0;
    const std::vector<double> wtx = diag_of(prhs[6], (size_t)(s.N + s.cp + s.cs));
    const std::vector<double> wrx = diag_of(prhs[11], (size_t)(s.N + s.tail_rx));
    const unsigned long long seed = nrhs > 14 ? (unsigned long long)mxGetScalar(prhs[14]) : 0ull;
    if (!sweep) {
        const int L = (int)mxGetNumberOfElements(prhs[7]);
        const std::vector<double> chan = interleave(prhs[7], (size_t)L, 1, 0);
        const double snr = mxGetScalar(prhs[8]);
        long long be = 0, bt = 0, se = 0, st = 0;
        check(wofdm_ber_run(handle(), &s, wtx.data(), wrx.data(), chan.data(), L, 1, &snr, 1, ensemble, seed, 0,
                            (int64_t*)&be, (int64_t*)&bt, (int64_t*)&se, (int64_t*)&st), "wofdm_ber_run");
        plhs[0] = mxCreateDoubleScalar(bt ? (double)be / (double)bt : 0.0);   // ber, as [~, ber] = biterr(...) at :272
        return;
    }
    // channels(channelIndex, :) -> column-major L x C complex
    const size_t C = mxGetM(prhs[7]), L = mxGetN(prhs[7]);
    std::vector<double> chan(2 * L * C);
    for (size_t c = 0; c < C; ++c) {
        const std::vector<double> row = interleave(prhs[7], L, C, c);
        memcpy(chan.data() + 2 * L * c, row.data(), 2 * L * sizeof(double));
    }
    const size_t n_snr = mxGetNumberOfElements(prhs[8]);
    const double* snr = mxGetPr(prhs[8]);
    std::vector<long long> be(n_snr), bt(n_snr), se(n_snr), st(n_snr);
    check(wofdm_ber_run(handle(), &s, wtx.data(), wrx.data(), chan.data(), (int)L, (int)C, snr, (int)n_snr, ensemble, seed, 0,
                        (int64_t*)be.data(), (int64_t*)bt.data(), (int64_t*)se.data(), (int64_t*)st.data()), "wofdm_ber_run");
    plhs[0] = mxCreateDoubleMatrix((mwSize)n_snr, 1, mxREAL);
    double* out = mxGetPr(plhs[0]);
    for (size_t i = 0; i < n_snr; ++i) out[i] = bt[i] ? (double)be[i] / (double)bt[i] : 0.0;
}

static void do_calculate_interference(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
    if (nrhs < 8) mexErrMsgIdAndTxt("wofdm:nargin", "calculate_interference needs 8 arguments");
    (void)nlhs;
    char name[16];
    if (!mxIsChar(prhs[1]) || mxGetString(prhs[1], name, sizeof(name))) mexErrMsgIdAndTxt("wofdm:type", "typeOFDM must be a string");
    wofdm_sys_t s;
    memset(&s, 0, sizeof(s));
    s.bits = 4; s.S = 2; s.precision = 1;
    if (wofdm_params_from_name(name, (int)mxGetScalar(prhs[4]), (int)mxGetScalar(prhs[0]), (int)mxGetScalar(prhs[5]),
                               (int)mxGetScalar(prhs[6]), &s) != WOFDM_OK)
        mexErrMsgIdAndTxt("wofdm:type", "unknown typeOFDM '%s'", name);
    const std::vector<double> wtx = diag_of(prhs[2], (size_t)(s.N + s.cp + s.cs));
    const std::vector<double> wrx = diag_of(prhs[3], (size_t)(s.N + s.tail_rx));
    // channels: rows = realisations (vehA200channel2(channelIndex, :)); the reference uses their mean (:196)
    const size_t C = mxGetM(prhs[7]), L = mxGetN(prhs[7]);
    std::vector<double> mean(2 * L, 0.0);
    for (size_t c = 0; c < C; ++c) {
        const std::vector<double> row = interleave(prhs[7], L, C, c);
        for (size_t i = 0; i < 2 * L; ++i) mean[i] += row[i] / (double)C;
    }
    const int mode = nrhs > 8 ? (int)mxGetScalar(prhs[8]) : 0;
    double P = 0.0;
    check(wofdm_interf_power_scalar(handle(), &s, wtx.data(), wrx.data(), mean.data(), (int)L, 1, mode, &P),
          "wofdm_interf_power_scalar");
    plhs[0] = mxCreateDoubleScalar(P);
}

extern "C" void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
    char cmd[32];
    if (nrhs < 1 || !mxIsChar(prhs[0]) || mxGetString(prhs[0], cmd, sizeof(cmd)))
        mexErrMsgIdAndTxt("wofdm:usage", "first argument must be 'run_simulation', 'run_simulation_sweep' or 'calculate_interference'");
    if (!strcmp(cmd, "run_simulation")) do_run_simulation(nlhs, plhs, nrhs - 1, prhs + 1, false);
    else if (!strcmp(cmd, "run_simulation_sweep")) do_run_simulation(nlhs, plhs, nrhs - 1, prhs + 1, true);
    else if (!strcmp(cmd, "calculate_interference")) do_calculate_interference(nlhs, plhs, nrhs - 1, prhs + 1);
    else mexErrMsgIdAndTxt("wofdm:usage", "unknown command '%s'", cmd);
}
