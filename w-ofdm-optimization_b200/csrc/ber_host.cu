// ber_host.cu -- host side of K1: variant selection, device tables, plans, and the
// wofdm_ber_* entry points of include/wofdm.h.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <new>

#include "host_common.h"
#include "mask_kernel.cuh"
#include "mask_gemm.h"
#include "ber_tconv2.cuh"

using namespace wofdm;

struct PlanDev {
    void *d_wtx = nullptr, *d_wrx = nullptr, *d_tw = nullptr, *d_chan = nullptr, *d_snr = nullptr;
    unsigned long long* d_cnt = nullptr;
    void* d_scratch = nullptr;
    size_t scratch_bytes = 0;
    int blocks_per_sm = 0;
    long long max_ctas = 0;
    bool pending = false;
    cudaStream_t last_stream = nullptr;
};

struct wofdm_ber_plan_s {
    wofdm_ctx* ctx = nullptr;
    wofdm_sys_t sys{};
    int L = 0, C = 0, n_snr = 0;
    const BerVariant* var = nullptr;
    BerSmem lay{};
    int chunk = 0, use_global = 0;
    int flat_tx = 0, flat_rx = 0;  // bit v = window pair v
    int n_var = 1;                 // window pairs of the plan
    bool fused = false;            // all of them in ONE launch (second-generation tensor-core kernel); else one launch each
    bool transient = false;    // buffers live in the devices' arenas (one-shot plan of wofdm_ber_run*): nothing to free
    std::vector<PlanDev> devs;
};

namespace {

struct Choice {
    const BerVariant* var = nullptr;
    BerSmem lay{};
    int chunk = 0, use_global = 0;
};

// tensor-core kernel: windows that are one value between their tails (BerParams::flat_tx / flat_rx)
// (n_var window pairs back to back: bit v of flat_tx / flat_rx = pair v)
void fill_flat(BerParams& prm, const wofdm_sys_t& s, const double* win_tx, const double* win_rx, int n_var = 1) {
    const int n_tx = s.N + s.cp + s.cs, n_wr = s.N + s.tail_rx;
    prm.flat_tx = prm.flat_rx = 0;
    for (int v = 0; v < n_var; ++v) {
        const double* wt = win_tx + (size_t)v * n_tx;
        const double* wr = win_rx + (size_t)v * n_wr;
        bool ft = s.cp >= s.tail_tx && s.cs >= s.tail_tx && wt[s.tail_tx] != 0.0;
        for (int i = s.tail_tx; ft && i < n_tx - s.tail_tx; ++i) ft = wt[i] == wt[s.tail_tx];
        bool fr = wr[s.tail_rx] != 0.0;
        for (int i = s.tail_rx; fr && i < s.N; ++i) fr = wr[i] == wr[s.tail_rx];
        prm.flat_tx |= (ft ? 1 : 0) << v;
        prm.flat_rx |= (fr ? 1 : 0) << v;
    }
}

// noise numbering of a frame shared by a cluster (BerParams::split)
void fill_split(BerParams& prm, const BerVariant& v) {
    prm.split = v.CL > 1 ? prm.S * prm.stride / v.CL : 0;
    prm.split_nt = v.CL > 1 ? v.NT : 0;
    prm.n48 = v.gen == 2 ? 1 : 0;      // 48-bit noise draws in receiver layout (ber_kernel.cuh: noise_draw48)
}

// win_tx (may be NULL): the circular-interior kernels need it flat between the tails; no_circ excludes them
int choose_variant(wofdm_ctx* h, const wofdm_sys_t& s, int L, bool verify, bool force_staged, size_t smem_cap,
                   Choice* out, const double* win_tx = nullptr, bool no_circ = false, bool want_txs = false,
                   bool no_tconv = false, int nvar = 1, bool want_txy = false) {
    const int stride = s.N + s.cp + s.cs - s.tail_tx;
    const int sec = s.S * stride;
    const bool fp64 = s.precision == 1;
    Choice best;
    bool flat = win_tx != nullptr && !no_circ && !getenv("WOFDM_NO_CIRC");
    if (flat)
        for (int i = s.tail_tx; i < s.N + s.cp + s.cs - s.tail_tx; ++i)
            if (win_tx[i] != win_tx[s.tail_tx]) { flat = false; break; }
    // WOFDM_VARIANT=<substring> restricts the tuned candidates (kernel tuning aid, e.g. "_b2")
    const char* want = getenv("WOFDM_VARIANT");
    if (no_tconv || getenv("WOFDM_NO_TCONV")) no_tconv = true;
    const int tconv_gen = getenv("WOFDM_TCONV_GEN") ? atoi(getenv("WOFDM_TCONV_GEN")) : 2;   // 1: ber_tconv.cuh, 2: ber_tconv2.cuh
    if (!fp64 && !force_staged) {
        for (const auto& v : h->variants) {
            if (v.TC == 0 || v.fp64 || v.verify != verify || v.N != s.N) continue;
            if (want && !strstr(v.name, want)) continue;
            if (v.ntile > 0) {
                // channel convolution on the tensor cores (ber_tconv.cuh): first choice wherever it applies -- one Tx pass,
                // L <= 21, prefix / suffix / tails inside the outer register rows, power sums and the zeros behind the
                // stream inside the frame's tiles.  Noise numbering: draw = position (chunk 0).  Takes tx_stream too.
                // CL CTAs share a frame: S/CL consecutive symbols each, in one Tx pass of the CTA
                if (s.S % v.CL != 0) continue;
                const int tpf = s.N / 16, S_cta = s.S / v.CL, sec_cta = sec / v.CL, body = s.tail_tx + sec_cta;
                if (no_tconv || S_cta > v.NT / tpf || L > v.LB) continue;
                if (v.gen != tconv_gen) continue;
                if (v.txy != want_txy) continue;                     // (the mask product's consumers serve the channel-mask variant only)
                if (v.CL > 1 && want_txs) continue;                  // (the masked Tx stream is a one-CTA-per-frame feature)
                // second generation: a receiver thread holds at most N48_MAXLEV noise samples besides its 16 FFT rows
                if (v.gen == 2 && (stride - s.N + (s.noise_norm == 1 ? s.tail_tx + L - 1 : 0) + tpf - 1) / tpf > (v.LB > TCV_LB ? N48_MAXLEV : N48_MAXLEV_SHORT)) continue;
                if (s.cp > 2 * tpf || s.cs > 2 * tpf || s.tail_tx > 2 * tpf || s.tail_rx / 2 > tpf || s.shift > tpf) continue;
                const int need = std::max(s.noise_norm == 1 ? body + L - 1 : sec_cta, body + (v.gen == 2 ? tconv2_zero(v.LB) : TCV_ZERO));
                if (need > 512 * v.ntile || sec_cta <= 512 * (v.ntile - 2)) continue;   // (the kernel range-checks its last two tiles only)
                if (nvar > 1 && v.gen != 2) continue;                // (several window pairs per launch: ber_tconv2.cuh only)
                const BerSmem lay = v.layout(S_cta, stride, s.tail_tx, s.tail_rx, L, v.gen == 2 ? nvar : 0, 0);
                if (lay.bytes > smem_cap) continue;
                // fewest taps of history first (MMAs per tile), then fewest tiles
                if (!best.var || best.var->ntile == 0 || v.LB < best.var->LB || (v.LB == best.var->LB && v.ntile < best.var->ntile)) {
                    best.var = &v; best.lay = lay; best.chunk = 0;
                }
                continue;
            }
            if (best.var && best.var->ntile > 0) continue;
            if (v.txs != want_txs) continue;                 // the tx_stream instantiations serve the channel-mask variant only
            // CL CTAs share a frame: S/CL consecutive symbols each, all of them in one pass of the CTA
            if (s.S % v.CL != 0 || (v.CL > 1 && s.S / v.CL > v.NT / (v.N / 16))) continue;
            const int S_cta = s.S / v.CL, sec_cta = sec / v.CL;
            const int chunk = ((sec_cta + v.NT - 1) / v.NT) | 1;
            if (chunk > v.TC || L > v.LB) continue;
            // the tuned kernels only look at the outer register rows for the prefix / suffix / heads / overlap-add
            // (ber_kernel.cuh, ER): rows are N/16 samples wide
            const int tpf = s.N / 16;
            if (s.cp > 2 * tpf || s.cs > 2 * tpf || s.tail_tx > 2 * tpf || s.tail_rx / 2 > tpf || s.shift > tpf) continue;
            if (v.full && !(chunk == v.TC && v.NT * v.TC == sec_cta)) continue;
            // circular interior: flat window, the tail_tx + L - 1 edge outputs of a symbol two per thread, one Tx pass
            if (v.circ && !(flat && s.tail_tx + L - 1 <= 2 * tpf && s.S <= v.NT / tpf && s.cp + 2 * tpf >= s.tail_tx + L - 1)) continue;
            const BerSmem lay = v.layout(S_cta, stride, s.tail_tx, s.tail_rx, L, chunk, 0);
            if (lay.bytes > smem_cap) continue;
            // measured order: the exact-fit direct kernel, then the circular-interior kernels
            // (+8 % over a direct kernel whose chunk does not fit exactly, -2..+1 % against the exact fit), then the rest
            auto rank = [](const BerVariant& x) { return (x.full && !x.circ) ? 0 : x.circ ? 1 : 2; };
            const bool better = !best.var || rank(v) < rank(*best.var) ||
                                (rank(v) == rank(*best.var) && (v.TC < best.var->TC || (v.TC == best.var->TC && v.full && !best.var->full)));
            if (better) { best.var = &v; best.lay = lay; best.chunk = chunk; }
        }
    }
    if (!best.var) {
        for (const auto& v : h->variants) {
            if (v.TC != 0 || v.fp64 != fp64 || v.verify != verify || v.N != s.N) continue;
            best.var = &v;
            best.chunk = ((sec + 255) / 256) | 1;   // noise block B of the staged policy (any odd number)
            best.lay = v.layout(s.S, stride, s.tail_tx, s.tail_rx, L, 0, 0);
            best.use_global = 0;
            if (best.lay.bytes > smem_cap) {
                best.use_global = 1;
                best.lay = v.layout(s.S, stride, s.tail_tx, s.tail_rx, L, 0, 1);
                if (best.lay.bytes > smem_cap) return fail(h, WOFDM_EUNSUPPORTED, "frame does not fit the staged kernel's shared memory");
            }
            break;
        }
    }
    if (!best.var) return fail(h, WOFDM_EUNSUPPORTED, "no compiled kernel variant for this N / precision");
    *out = best;
    return WOFDM_OK;
}

template <typename T> void cast_vec(const std::vector<double>& src, std::vector<unsigned char>& dst) {
    dst.resize(src.size() * sizeof(T));
    T* p = reinterpret_cast<T*>(dst.data());
    for (size_t i = 0; i < src.size(); ++i) p[i] = (T)src[i];
}
void cast_any(bool fp64, const std::vector<double>& src, std::vector<unsigned char>& dst) {
    if (fp64) cast_vec<double>(src, dst); else cast_vec<float>(src, dst);
}

// host tables shared by plans and the verify path
struct HostTables {
    std::vector<unsigned char> wtx, wrx, tw;
};

// unit_peak (tensor-core kernels): the Tx window is divided by its largest magnitude.  Those kernels stage the stream and
// the taps as fp16 pairs behind fixed power-of-two scales (ber_tconv.cuh: TCV_XSCALE, TCV_HSCALE), so their inputs must
// have a known range; the chain itself does not care -- a common factor of the Tx signal or of a channel's taps scales
// r, the noise follows the measured signal power (wofdm_simulation.py:135-138) and the pilot equaliser divides it out (:223-232).
void build_tables(const wofdm_sys_t& s, const double* win_tx, const double* win_rx, HostTables& t, bool unit_peak = false) {
    const bool fp64 = s.precision == 1;
    const int n_tx = s.N + s.cp + s.cs;
    double k = qam_scale(s) / (double)s.N;   // IDFT 1/N (transmitter.py:58) and constellation scale
    if (unit_peak) {
        double mx = 0.0;
        for (int i = 0; i < n_tx; ++i) mx = std::max(mx, std::fabs(win_tx[i]));
        if (mx > 0.0 && std::isfinite(mx)) k /= mx;
    }
    std::vector<double> a(n_tx), b(s.N + s.tail_rx);
    for (int i = 0; i < n_tx; ++i) a[i] = win_tx[i] * k;
    for (int i = 0; i < s.N + s.tail_rx; ++i) b[i] = win_rx[i];
    cast_any(fp64, a, t.wtx);
    cast_any(fp64, b, t.wrx);
    cast_any(fp64, build_twiddles(s.N), t.tw);
}

// channel matrix L x C column-major complex double -> [C][L] V2<T> (same memory order, cast only).  unit_peak (tensor-core
// kernels, see build_tables): every channel's taps are divided by their largest component magnitude.
void cast_chan(bool fp64, const double* chan, size_t n_complex, std::vector<unsigned char>& dst, int L = 0, bool unit_peak = false) {
    std::vector<double> v(chan, chan + 2 * n_complex);
    if (unit_peak && L > 0) {
        for (size_t c = 0; c + (size_t)L <= n_complex; c += (size_t)L) {
            double mx = 0.0;
            for (int i = 0; i < 2 * L; ++i) mx = std::max(mx, std::fabs(v[2 * c + i]));
            if (mx > 0.0 && std::isfinite(mx)) {
                const double inv = 1.0 / mx;
                for (int i = 0; i < 2 * L; ++i) v[2 * c + i] *= inv;
            }
        }
    }
    cast_any(fp64, v, dst);
}

void fill_sys(BerParams& p, const wofdm_sys_t& s, int L) {
    memset(&p, 0, sizeof(p));
    p.N = s.N; p.cp = s.cp; p.cs = s.cs; p.tail_tx = s.tail_tx; p.tail_rx = s.tail_rx; p.rm = s.rm;
    p.shift = s.shift; p.bits = s.bits; p.S = s.S;
    p.n_tx = s.N + s.cp + s.cs;
    p.stride = p.n_tx - s.tail_tx;
    p.L = L;
    p.noise_norm = s.noise_norm; p.constellation = s.constellation;
    p.guard = s.guard;
    p.noise_len = noise_len(s, L);
    p.qscale = qam_scale(s);
}

// max_ctas: CTAs of this variant that can be resident on the device at once (a multiple of the cluster size)
int prepare_kernel(wofdm_ctx* h, const BerVariant& v, size_t smem, int sm_count, int* blocks_per_sm, long long* max_ctas) {
    WOFDM_CUDA(h, cudaFuncSetAttribute(v.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int nb = 0;
    WOFDM_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, v.fn, v.launch_threads, smem));
    if (v.ntile > 0) {
        // the occupancy calculator counts a kernel that allocates tensor memory as one CTA per SM; these kernels take
        // 256 of the 512 columns (N = 256; all of them at N = 512), so registers, shared memory and tensor memory together decide
        cudaFuncAttributes fa;
        int dev = 0, smem_sm = 0, regs_sm = 0;
        WOFDM_CUDA(h, cudaFuncGetAttributes(&fa, v.fn));
        WOFDM_CUDA(h, cudaGetDevice(&dev));
        WOFDM_CUDA(h, cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev));
        WOFDM_CUDA(h, cudaDeviceGetAttribute(&regs_sm, cudaDevAttrMaxRegistersPerMultiprocessor, dev));
        const int by_smem = (int)((size_t)smem_sm / (smem + 1024)), by_regs = regs_sm / (((fa.numRegs + 7) & ~7) * v.launch_threads);
        const int by_tmem = 512 / (int)tconv_tmem_cols(v.ntile);
        nb = std::max(nb, std::min(by_smem, std::min(by_regs, by_tmem)));
    }
    if (const char* f = getenv("WOFDM_FORCE_CTAS_PER_SM")) nb = std::max(1, atoi(f));    // tuning aid
    if (getenv("WOFDM_DEBUG")) fprintf(stderr, "[wofdm] %s: smem %zu B, %d CTAs/SM\n", v.name, smem, nb);
    if (nb < 1) return fail(h, WOFDM_EUNSUPPORTED, "kernel variant cannot be resident on this device");
    *blocks_per_sm = nb;
    long long cap = (long long)nb * sm_count;
    if (v.CL > 1) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(v.CL * sm_count); cfg.blockDim = dim3(v.launch_threads); cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = v.CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        int ncl = 0;
        WOFDM_CUDA(h, cudaOccupancyMaxActiveClusters(&ncl, v.fn, &cfg));
        if (ncl < 1) return fail(h, WOFDM_EUNSUPPORTED, "no thread-block cluster of this variant fits the device");
        // (tensor-memory kernels: the cluster occupancy query, too, counts one CTA per SM; nb above is the real residency.
        //  A grid larger than what is resident is still correct -- frames are taken round-robin by cluster index.)
        cap = std::max<long long>((long long)ncl * v.CL, v.ntile > 0 ? ((long long)nb * sm_count / v.CL) * v.CL : 0);
    }
    if (max_ctas) *max_ctas = cap;
    return WOFDM_OK;
}

size_t elem_bytes(const wofdm_sys_t& s) { return s.precision == 1 ? sizeof(double2) : sizeof(float2); }

}  // namespace

// transient = true: device buffers come out of the per-device arena (no cudaMalloc / cudaFree per call); such a plan
// must be destroyed before the arena is used again (wofdm_ber_run_shard does exactly that)
static int plan_create_impl(wofdm_handle h, const wofdm_sys_t* sys, const double* win_tx, const double* win_rx, int n_var,
                            const double* chan, int L, int C, const double* snr_db, int n_snr,
                            wofdm_ber_plan* out, bool transient) {
    NvtxRange nvtx_("wofdm_ber_plan_create");
    if (!h) return WOFDM_EINVAL;
    if (!out) return fail(h, WOFDM_EINVAL, "plan out pointer is NULL");
    *out = nullptr;
    int rc = validate_sys(h, sys, L);
    if (rc) return rc;
    if (!win_tx || !win_rx || !chan || !snr_db) return fail(h, WOFDM_EINVAL, "NULL input buffer");
    if (C < 1 || n_snr < 1) return fail(h, WOFDM_EINVAL, "C and n_snr must be >= 1");
    if (n_var < 1 || n_var > WOFDM_MAX_VARIANTS) return fail(h, WOFDM_EINVAL, "n_var must be in [1, WOFDM_MAX_VARIANTS]");
    const bool fp64 = sys->precision == 1;

    wofdm_ber_plan_s* p = new (std::nothrow) wofdm_ber_plan_s();
    if (!p) return fail(h, WOFDM_ENOMEM, "host allocation failed");
    p->ctx = h; p->sys = *sys; p->L = L; p->C = C; p->n_snr = n_snr; p->transient = transient; p->n_var = n_var;
    Choice ch;
    size_t cap = h->devs[0].smem_optin;
    for (auto& d : h->devs) cap = std::min(cap, d.smem_optin);
    // several window pairs: one launch for all of them where the second-generation tensor-core kernel applies and holds
    // their tables; otherwise the pairs are launched one after the other on the same draws (same counters either way)
    rc = choose_variant(h, *sys, L, false, false, cap, &ch, win_tx, false, false, false, n_var);
    p->fused = rc == WOFDM_OK && n_var > 1 && ch.var->gen == 2;
    if (n_var > 1 && !p->fused) rc = choose_variant(h, *sys, L, false, false, cap, &ch, win_tx, /*no_circ=*/true);
    if (rc) { delete p; return rc; }
    p->var = ch.var; p->lay = ch.lay; p->chunk = ch.chunk; p->use_global = ch.use_global;
    {
        BerParams fp;
        fill_flat(fp, *sys, win_tx, win_rx, n_var);
        p->flat_tx = fp.flat_tx; p->flat_rx = fp.flat_rx;
    }

    HostTables t;
    const bool unit_peak = p->var->ntile > 0;
    const int n_tx = sys->N + sys->cp + sys->cs, n_wr = sys->N + sys->tail_rx;
    for (int v = 0; v < n_var; ++v) {              // tables of the window pairs back to back
        HostTables tv;
        build_tables(*sys, win_tx + (size_t)v * n_tx, win_rx + (size_t)v * n_wr, tv, unit_peak);
        t.wtx.insert(t.wtx.end(), tv.wtx.begin(), tv.wtx.end());
        t.wrx.insert(t.wrx.end(), tv.wrx.begin(), tv.wrx.end());
        if (v == 0) t.tw = tv.tw;
    }
    std::vector<unsigned char> hchan, hsnr;
    cast_chan(fp64, chan, (size_t)L * C, hchan, L, unit_peak);
    std::vector<double> lin(n_snr);
    for (int i = 0; i < n_snr; ++i) lin[i] = std::pow(10.0, -0.1 * snr_db[i]);   // wofdm_simulation.py:138
    cast_any(fp64, lin, hsnr);

    p->devs.resize(h->devs.size());
    for (size_t i = 0; i < h->devs.size(); ++i) {
        DeviceCtx& d = h->devs[i];
        PlanDev& pd = p->devs[i];
        auto dev_alloc = [&](void** dst, size_t bytes) -> cudaError_t {
            if (!transient) return cudaMalloc(dst, std::max<size_t>(bytes, 16));
            *dst = arena_take(d, std::max<size_t>(bytes, 16));
            return *dst ? cudaSuccess : cudaErrorMemoryAllocation;
        };
        auto up = [&](void** dst, const std::vector<unsigned char>& src) -> cudaError_t {
            cudaError_t e = dev_alloc(dst, src.size());
            if (e != cudaSuccess) return e;
            return cudaMemcpyAsync(*dst, src.data(), src.size(), cudaMemcpyHostToDevice, d.stream);
        };
        cudaError_t e = cudaSetDevice(d.dev);
        if (e == cudaSuccess) {
            rc = prepare_kernel(h, *p->var, p->lay.bytes, d.sm_count, &pd.blocks_per_sm, &pd.max_ctas);
            if (rc) { const std::string msg = h->err; wofdm_ber_plan_destroy(p); h->err = msg; return rc; }
        }
        // staged policy with frame buffers in global memory: two of them per resident CTA
        const size_t scratch_elems = (size_t)p->lay.pad + sys->tail_tx + (size_t)sys->S * (sys->N + sys->cp + sys->cs - sys->tail_tx) + 64;
        pd.scratch_bytes = p->use_global ? (size_t)pd.blocks_per_sm * d.sm_count * 2 * scratch_elems * elem_bytes(*sys) : 0;
        if (e == cudaSuccess && transient) {
            rc = arena_reserve(h, d, t.wtx.size() + t.wrx.size() + t.tw.size() + hchan.size() + hsnr.size() +
                                         (size_t)n_var * n_snr * 16 + pd.scratch_bytes);
            if (rc) { const std::string msg = h->err; wofdm_ber_plan_destroy(p); h->err = msg; return rc; }
        }
        if (e == cudaSuccess) e = up(&pd.d_wtx, t.wtx);
        if (e == cudaSuccess) e = up(&pd.d_wrx, t.wrx);
        if (e == cudaSuccess) e = up(&pd.d_tw, t.tw);
        if (e == cudaSuccess) e = up(&pd.d_chan, hchan);
        if (e == cudaSuccess) e = up(&pd.d_snr, hsnr);
        if (e == cudaSuccess) e = dev_alloc(reinterpret_cast<void**>(&pd.d_cnt), (size_t)n_var * n_snr * 2 * sizeof(unsigned long long));
        if (e == cudaSuccess) e = cudaStreamSynchronize(d.stream);   // host vectors go out of scope
        if (e == cudaSuccess && p->use_global) e = dev_alloc(&pd.d_scratch, pd.scratch_bytes);
        if (e != cudaSuccess) {
            std::string msg = std::string("plan upload: ") + cudaGetErrorString(e);
            wofdm_ber_plan_destroy(p);
            return fail(h, e == cudaErrorMemoryAllocation ? WOFDM_ENOMEM : WOFDM_ECUDA, msg);
        }
    }
    *out = p;
    return WOFDM_OK;
}

extern "C" {

int wofdm_ber_plan_create(wofdm_handle h, const wofdm_sys_t* sys, const double* win_tx, const double* win_rx,
                          const double* chan, int L, int C, const double* snr_db, int n_snr,
                          wofdm_ber_plan* out) {
    return plan_create_impl(h, sys, win_tx, win_rx, 1, chan, L, C, snr_db, n_snr, out, false);
}

int wofdm_ber_plan_create_multi(wofdm_handle h, const wofdm_sys_t* sys, const double* win_tx, const double* win_rx, int n_var,
                                const double* chan, int L, int C, const double* snr_db, int n_snr, wofdm_ber_plan* out) {
    return plan_create_impl(h, sys, win_tx, win_rx, n_var, chan, L, C, snr_db, n_snr, out, false);
}

int wofdm_ber_plan_variants(wofdm_ber_plan p, int* fused) {
    if (!p) return WOFDM_EINVAL;
    if (fused) *fused = p->fused ? 1 : 0;
    return p->n_var;
}

int wofdm_ber_plan_launch(wofdm_ber_plan p, int slot, int64_t ensemble, uint64_t seed, uint32_t variant,
                          int shard_index, int shard_count, void* stream, void** d_counters) {
    NvtxRange nvtx_("wofdm_ber_plan_launch");
    if (!p) return WOFDM_EINVAL;
    wofdm_ctx* h = p->ctx;
    if (slot < 0 || slot >= (int)p->devs.size()) return fail(h, WOFDM_EINVAL, "device slot out of range");
    if (ensemble < 1) return fail(h, WOFDM_EINVAL, "ensemble must be >= 1");
    if (shard_count < 1 || shard_index < 0 || shard_index >= shard_count) return fail(h, WOFDM_EINVAL, "bad shard");
    if (variant > 0xfffffff0u - WOFDM_MAX_VARIANTS) return fail(h, WOFDM_EINVAL, "variant too large");
    DeviceCtx& d = h->devs[slot];
    PlanDev& pd = p->devs[slot];
    cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : d.stream;
    WOFDM_CUDA(h, cudaSetDevice(d.dev));

    const long long total = (long long)p->n_snr * p->C * ensemble;
    const long long mine = total > shard_index ? (total - shard_index + shard_count - 1) / shard_count : 0;
    BerParams prm;
    fill_sys(prm, p->sys, p->L);
    prm.chunk = p->chunk; prm.use_global = p->use_global;
    prm.flat_tx = p->flat_tx; prm.flat_rx = p->flat_rx;
    fill_split(prm, *p->var);
    prm.win_tx = pd.d_wtx; prm.win_rx = pd.d_wrx; prm.tw = pd.d_tw; prm.chan = pd.d_chan; prm.snr_lin = pd.d_snr;
    prm.C = p->C; prm.n_snr = p->n_snr; prm.ensemble = ensemble;
    prm.seed = seed; prm.variant = variant;
    philox_round_keys(seed, prm.rk);
    prm.frame_begin = shard_index; prm.frame_step = shard_count; prm.n_frames = mine;
    prm.counters = pd.d_cnt;
    prm.scratch = pd.d_scratch;
    prm.scratch_elems = p->use_global ? (long long)(pd.scratch_bytes / ((size_t)pd.blocks_per_sm * d.sm_count * 2 * elem_bytes(p->sys))) : 0;

    // a launch owns its slot's counters: one that is still pending on ANOTHER stream must finish before they are zeroed
    if (pd.pending && pd.last_stream != st) WOFDM_CUDA(h, cudaStreamSynchronize(pd.last_stream));
    WOFDM_CUDA(h, cudaMemsetAsync(pd.d_cnt, 0, (size_t)p->n_var * p->n_snr * 2 * sizeof(unsigned long long), st));
    if (mine > 0) {
        long long cap = pd.max_ctas;
        if (const char* lim = getenv("WOFDM_MAX_CTAS_PER_SM"))            // tuning aid: occupancy sensitivity
            cap = std::min<long long>(cap, (long long)std::max(1, atoi(lim)) * d.sm_count);
        const int cl = p->var->CL;
        const int grid = (int)std::min<long long>(mine, std::max<long long>(cap / cl, 1)) * cl;   // CTAs = frames in flight x cluster size
        if (p->fused || p->n_var == 1) {
            prm.nvar = p->n_var;
            WOFDM_CUDA(h, p->var->launch(prm, grid, p->lay.bytes, st));
            h->launches += 1;
        } else {
            // one launch per window pair: pair v = a single-pair launch with variant + v on tables / counters v
            const size_t eb = p->sys.precision == 1 ? sizeof(double) : sizeof(float);
            const int n_tx = p->sys.N + p->sys.cp + p->sys.cs, n_wr = p->sys.N + p->sys.tail_rx;
            for (int v = 0; v < p->n_var; ++v) {
                BerParams pv = prm;
                pv.nvar = 1;
                pv.win_tx = static_cast<const char*>(pd.d_wtx) + (size_t)v * n_tx * eb;
                pv.win_rx = static_cast<const char*>(pd.d_wrx) + (size_t)v * n_wr * eb;
                pv.flat_tx = (p->flat_tx >> v) & 1; pv.flat_rx = (p->flat_rx >> v) & 1;
                pv.variant = variant + (uint32_t)v;
                pv.counters = pd.d_cnt + (size_t)v * p->n_snr * 2;
                WOFDM_CUDA(h, p->var->launch(pv, grid, p->lay.bytes, st));
                h->launches += 1;
            }
        }
    }
    pd.pending = true;
    pd.last_stream = st;
    if (d_counters) *d_counters = pd.d_cnt;
    return WOFDM_OK;
}

int wofdm_ber_plan_read(wofdm_ber_plan p, int64_t* bit_err, int64_t* sym_err) {
    if (!p) return WOFDM_EINVAL;
    wofdm_ctx* h = p->ctx;
    if (!bit_err || !sym_err) return fail(h, WOFDM_EINVAL, "NULL output buffer");
    const int n_out = p->n_var * p->n_snr;       // [n_var][n_snr]
    for (int i = 0; i < n_out; ++i) { bit_err[i] = 0; sym_err[i] = 0; }
    std::vector<unsigned long long> tmp((size_t)n_out * 2);
    for (size_t s = 0; s < p->devs.size(); ++s) {
        PlanDev& pd = p->devs[s];
        if (!pd.pending) continue;
        WOFDM_CUDA(h, cudaSetDevice(h->devs[s].dev));
        WOFDM_CUDA(h, cudaMemcpyAsync(tmp.data(), pd.d_cnt, tmp.size() * sizeof(unsigned long long),
                                      cudaMemcpyDeviceToHost, pd.last_stream));
        WOFDM_CUDA(h, cudaStreamSynchronize(pd.last_stream));
        for (int i = 0; i < n_out; ++i) { bit_err[i] += (int64_t)tmp[2 * i]; sym_err[i] += (int64_t)tmp[2 * i + 1]; }
        pd.pending = false;
    }
    return WOFDM_OK;
}

const char* wofdm_ber_plan_kernel(wofdm_ber_plan p) { return (p && p->var) ? p->var->name : ""; }

int wofdm_ber_plan_destroy(wofdm_ber_plan p) {
    if (!p) return WOFDM_EINVAL;
    for (size_t s = 0; s < p->devs.size(); ++s) {
        PlanDev& pd = p->devs[s];
        cudaSetDevice(p->ctx->devs[s].dev);
        if (pd.pending && pd.last_stream) cudaStreamSynchronize(pd.last_stream);
        if (p->transient) continue;             // arena memory
        cudaFree(pd.d_wtx); cudaFree(pd.d_wrx); cudaFree(pd.d_tw); cudaFree(pd.d_chan); cudaFree(pd.d_snr);
        cudaFree(pd.d_cnt); cudaFree(pd.d_scratch);
    }
    cudaGetLastError();
    delete p;
    return WOFDM_OK;
}

int wofdm_ber_run_shard(wofdm_handle h, const wofdm_sys_t* sys, const double* win_tx, const double* win_rx,
                        const double* chan, int L, int C, const double* snr_db, int n_snr, int64_t ensemble,
                        uint64_t seed, uint32_t variant, int shard_index, int shard_count,
                        int64_t* bit_err, int64_t* bit_tot, int64_t* sym_err, int64_t* sym_tot) {
    return wofdm_ber_run_multi(h, sys, win_tx, win_rx, 1, chan, L, C, snr_db, n_snr, ensemble, seed, variant, shard_index,
                               shard_count, bit_err, bit_tot, sym_err, sym_tot);
}

int wofdm_ber_run_multi(wofdm_handle h, const wofdm_sys_t* sys, const double* win_tx, const double* win_rx, int n_var,
                        const double* chan, int L, int C, const double* snr_db, int n_snr, int64_t ensemble,
                        uint64_t seed, uint32_t variant, int shard_index, int shard_count,
                        int64_t* bit_err, int64_t* bit_tot, int64_t* sym_err, int64_t* sym_tot) {
    NvtxRange nvtx_("wofdm_ber_run_multi");
    if (!h) return WOFDM_EINVAL;
    if (!bit_err || !bit_tot || !sym_err || !sym_tot) return fail(h, WOFDM_EINVAL, "NULL output buffer");
    if (shard_count < 1 || shard_index < 0 || shard_index >= shard_count) return fail(h, WOFDM_EINVAL, "bad shard");
    if (ensemble < 1) return fail(h, WOFDM_EINVAL, "ensemble must be >= 1");
    wofdm_ber_plan p = nullptr;
    int rc = plan_create_impl(h, sys, win_tx, win_rx, n_var, chan, L, C, snr_db, n_snr, &p, true);
    if (rc) return rc;
    // the handle's devices split this shard's frames between them: device i takes the sub-shard
    // shard_index + shard_count*i of shard_count*ndev, so results do not depend on the GPU count
    const int nd = (int)h->devs.size();
    for (int i = 0; i < nd && rc == WOFDM_OK; ++i)
        rc = wofdm_ber_plan_launch(p, i, ensemble, seed, variant, shard_index + shard_count * i, shard_count * nd,
                                   nullptr, nullptr);
    if (rc == WOFDM_OK) rc = wofdm_ber_plan_read(p, bit_err, sym_err);
    wofdm_ber_plan_destroy(p);
    if (rc) return rc;
    // frames of this shard per SNR point: f = (snr*C + c)*ensemble + e, f = shard_index (mod shard_count)
    const long long per_snr = (long long)C * ensemble;
    for (int i = 0; i < n_snr; ++i) {
        const long long lo = (long long)i * per_snr, hi = lo + per_snr;   // [lo, hi)
        auto count_upto = [&](long long x) -> long long {                  // #{f < x : f = shard_index mod shard_count}
            return x > shard_index ? (x - shard_index + shard_count - 1) / shard_count : 0;
        };
        const long long frames = count_upto(hi) - count_upto(lo);
        sym_tot[i] = frames * (long long)(sys->N - 2 * sys->guard) * (sys->S - 1);   // active sub-carriers only
        bit_tot[i] = sym_tot[i] * sys->bits;
    }
    return WOFDM_OK;
}

int wofdm_ber_run(wofdm_handle h, const wofdm_sys_t* sys, const double* win_tx, const double* win_rx,
                  const double* chan, int L, int C, const double* snr_db, int n_snr, int64_t ensemble,
                  uint64_t seed, uint32_t variant,
                  int64_t* bit_err, int64_t* bit_tot, int64_t* sym_err, int64_t* sym_tot) {
    return wofdm_ber_run_shard(h, sys, win_tx, win_rx, chan, L, C, snr_db, n_snr, ensemble, seed, variant, 0, 1,
                               bit_err, bit_tot, sym_err, sym_tot);
}

int wofdm_ber_verify(wofdm_handle h, const wofdm_sys_t* sys, const double* win_tx, const double* win_rx,
                     const double* chan, int L, int F, const double* snr_db,
                     const int32_t* sym_idx, const double* noise, int variant_kernel,
                     double* eq_out, int32_t* dec_idx, int64_t* bit_err, int64_t* sym_err) {
    NvtxRange nvtx_("wofdm_ber_verify");
    if (!h) return WOFDM_EINVAL;
    int rc = validate_sys(h, sys, L);
    if (rc) return rc;
    if (!win_tx || !win_rx || !chan || !snr_db || !sym_idx || !noise || !eq_out || !dec_idx || !bit_err || !sym_err)
        return fail(h, WOFDM_EINVAL, "NULL buffer");
    if (F < 1) return fail(h, WOFDM_EINVAL, "F must be >= 1");
    const bool fp64 = sys->precision == 1;
    const int M = 1 << sys->bits;
    const size_t n_sym = (size_t)sys->N * sys->S * F;
    for (size_t i = 0; i < n_sym; ++i)
        if (sym_idx[i] < 0 || sym_idx[i] >= M) return fail(h, WOFDM_EINVAL, "sym_idx entry outside the constellation");

    DeviceCtx& d = h->devs[0];
    WOFDM_CUDA(h, cudaSetDevice(d.dev));
    Choice ch;
    rc = choose_variant(h, *sys, L, true, variant_kernel == 1, d.smem_optin, &ch, win_tx, variant_kernel == 2, false,
                        variant_kernel == 2 || variant_kernel == 3);
    if (rc) return rc;
    int nb = 0;
    long long max_ctas = 0;
    rc = prepare_kernel(h, *ch.var, ch.lay.bytes, d.sm_count, &nb, &max_ctas);
    if (rc) return rc;
    const int grid = (int)std::min<long long>(F, std::max<long long>(max_ctas / ch.var->CL, 1)) * ch.var->CL;

    HostTables t;
    build_tables(*sys, win_tx, win_rx, t, ch.var->ntile > 0);
    std::vector<unsigned char> hchan, hsnr;
    cast_chan(fp64, chan, (size_t)L * F, hchan, L, ch.var->ntile > 0);
    std::vector<double> lin(F);
    for (int i = 0; i < F; ++i) lin[i] = std::pow(10.0, -0.1 * snr_db[i]);
    cast_any(fp64, lin, hsnr);
    const size_t nlen = (size_t)noise_len(*sys, L);
    const size_t n_eq = (size_t)sys->N * (sys->S - 1) * F;
    const size_t scratch_elems = (size_t)ch.lay.pad + sys->tail_tx + (size_t)sys->S * (sys->N + sys->cp + sys->cs - sys->tail_tx) + 64;
    const size_t scratch_bytes = ch.use_global ? (size_t)grid * 2 * scratch_elems * elem_bytes(*sys) : 0;

    const size_t total = t.wtx.size() + t.wrx.size() + t.tw.size() + hchan.size() + hsnr.size() + n_sym * 4 +
                         nlen * F * 16 + n_eq * 16 + n_eq * 4 + (size_t)F * 16 + scratch_bytes;
    rc = arena_reserve(h, d, total);
    if (rc) return rc;
    auto put = [&](const void* src, size_t bytes, void** dst) -> cudaError_t {
        *dst = arena_take(d, std::max<size_t>(bytes, 16));
        if (!*dst) return cudaErrorMemoryAllocation;
        return src ? cudaMemcpyAsync(*dst, src, bytes, cudaMemcpyHostToDevice, d.stream) : cudaSuccess;
    };
    void *d_wtx, *d_wrx, *d_tw, *d_chan, *d_snr, *d_sym, *d_noise, *d_eq, *d_dec, *d_be, *d_se, *d_scr = nullptr;
    WOFDM_CUDA(h, put(t.wtx.data(), t.wtx.size(), &d_wtx));
    WOFDM_CUDA(h, put(t.wrx.data(), t.wrx.size(), &d_wrx));
    WOFDM_CUDA(h, put(t.tw.data(), t.tw.size(), &d_tw));
    WOFDM_CUDA(h, put(hchan.data(), hchan.size(), &d_chan));
    WOFDM_CUDA(h, put(hsnr.data(), hsnr.size(), &d_snr));
    WOFDM_CUDA(h, put(sym_idx, n_sym * 4, &d_sym));
    WOFDM_CUDA(h, put(noise, nlen * F * 16, &d_noise));
    WOFDM_CUDA(h, put(nullptr, n_eq * 16, &d_eq));
    WOFDM_CUDA(h, put(nullptr, n_eq * 4, &d_dec));
    WOFDM_CUDA(h, put(nullptr, (size_t)F * 8, &d_be));
    WOFDM_CUDA(h, put(nullptr, (size_t)F * 8, &d_se));
    if (scratch_bytes) WOFDM_CUDA(h, put(nullptr, scratch_bytes, &d_scr));
    WOFDM_CUDA(h, cudaMemsetAsync(d_be, 0, (size_t)F * 8, d.stream));
    WOFDM_CUDA(h, cudaMemsetAsync(d_se, 0, (size_t)F * 8, d.stream));

    BerParams prm;
    fill_sys(prm, *sys, L);
    prm.chunk = ch.chunk; prm.use_global = ch.use_global;
    fill_flat(prm, *sys, win_tx, win_rx);
    fill_split(prm, *ch.var);
    prm.win_tx = d_wtx; prm.win_rx = d_wrx; prm.tw = d_tw; prm.chan = d_chan; prm.snr_lin = d_snr;
    prm.C = F; prm.n_snr = F; prm.ensemble = 1;
    prm.frame_begin = 0; prm.frame_step = 1; prm.n_frames = F;
    prm.sym_idx = static_cast<const int32_t*>(d_sym);
    prm.noise_in = static_cast<const double2*>(d_noise);
    prm.eq_out = static_cast<double2*>(d_eq);
    prm.dec_out = static_cast<int32_t*>(d_dec);
    prm.bit_err_f = static_cast<long long*>(d_be);
    prm.sym_err_f = static_cast<long long*>(d_se);
    prm.scratch = d_scr; prm.scratch_elems = (long long)scratch_elems;
    WOFDM_CUDA(h, ch.var->launch(prm, grid, ch.lay.bytes, d.stream));
    h->launches += 1;
    WOFDM_CUDA(h, cudaMemcpyAsync(eq_out, d_eq, n_eq * 16, cudaMemcpyDeviceToHost, d.stream));
    WOFDM_CUDA(h, cudaMemcpyAsync(dec_idx, d_dec, n_eq * 4, cudaMemcpyDeviceToHost, d.stream));
    WOFDM_CUDA(h, cudaMemcpyAsync(bit_err, d_be, (size_t)F * 8, cudaMemcpyDeviceToHost, d.stream));
    WOFDM_CUDA(h, cudaMemcpyAsync(sym_err, d_se, (size_t)F * 8, cudaMemcpyDeviceToHost, d.stream));
    WOFDM_CUDA(h, cudaStreamSynchronize(d.stream));
    return WOFDM_OK;
}

int wofdm_ber_draws(wofdm_handle h, const wofdm_sys_t* sys, int L, uint64_t seed, uint32_t variant,
                    const int64_t* frame_ids, int F, int32_t* sym_idx, double* noise) {
    NvtxRange nvtx_("wofdm_ber_draws");
    if (!h) return WOFDM_EINVAL;
    int rc = validate_sys(h, sys, L);
    if (rc) return rc;
    if (!frame_ids || !sym_idx || !noise || F < 1) return fail(h, WOFDM_EINVAL, "bad buffer");
    DeviceCtx& d = h->devs[0];
    WOFDM_CUDA(h, cudaSetDevice(d.dev));
    const size_t nlen = (size_t)noise_len(*sys, L);
    const size_t n_sym = (size_t)sys->N * sys->S * F;
    rc = arena_reserve(h, d, (size_t)F * 8 + n_sym * 4 + nlen * F * 16);
    if (rc) return rc;
    long long* d_ids = static_cast<long long*>(arena_take(d, (size_t)F * 8));
    int32_t* d_sym = static_cast<int32_t*>(arena_take(d, n_sym * 4));
    double2* d_noise = static_cast<double2*>(arena_take(d, nlen * F * 16));
    if (!d_ids || !d_sym || !d_noise) return fail(h, WOFDM_ENOMEM, "arena exhausted");
    WOFDM_CUDA(h, cudaMemcpyAsync(d_ids, frame_ids, (size_t)F * 8, cudaMemcpyHostToDevice, d.stream));
    BerParams prm;
    fill_sys(prm, *sys, L);
    prm.seed = seed; prm.variant = variant;
    philox_round_keys(seed, prm.rk);
    {   // the noise block B is a property of the kernel variant production mode dispatches to
        Choice ch;
        rc = choose_variant(h, *sys, L, false, false, d.smem_optin, &ch);
        if (rc) return rc;
        prm.chunk = ch.chunk;
        fill_split(prm, *ch.var);
    }
    const dim3 gs(sys->S, F);
    switch (sys->N) {
        case 16: draws_sym_kernel<16><<<gs, 32, 0, d.stream>>>(prm, d_ids, d_sym); break;
        case 32: draws_sym_kernel<32><<<gs, 32, 0, d.stream>>>(prm, d_ids, d_sym); break;
        case 64: draws_sym_kernel<64><<<gs, 32, 0, d.stream>>>(prm, d_ids, d_sym); break;
        case 128: draws_sym_kernel<128><<<gs, 32, 0, d.stream>>>(prm, d_ids, d_sym); break;
        case 256: draws_sym_kernel<256><<<gs, 32, 0, d.stream>>>(prm, d_ids, d_sym); break;
        case 512: draws_sym_kernel<512><<<gs, 32, 0, d.stream>>>(prm, d_ids, d_sym); break;
        case 1024: draws_sym_kernel<1024><<<gs, 64, 0, d.stream>>>(prm, d_ids, d_sym); break;
        default: return fail(h, WOFDM_EUNSUPPORTED, "N");
    }
    WOFDM_CUDA(h, cudaGetLastError());
    const dim3 gn((unsigned)((nlen + 127) / 128), F);
    if (sys->precision == 1) draws_noise_kernel<double><<<gn, 128, 0, d.stream>>>(prm, d_ids, d_noise);
    else draws_noise_kernel<float><<<gn, 128, 0, d.stream>>>(prm, d_ids, d_noise);
    WOFDM_CUDA(h, cudaGetLastError());
    h->launches += 2;
    WOFDM_CUDA(h, cudaMemcpyAsync(sym_idx, d_sym, n_sym * 4, cudaMemcpyDeviceToHost, d.stream));
    WOFDM_CUDA(h, cudaMemcpyAsync(noise, d_noise, nlen * F * 16, cudaMemcpyDeviceToHost, d.stream));
    WOFDM_CUDA(h, cudaStreamSynchronize(d.stream));
    return WOFDM_OK;
}

// in-place forward FFT of n = 2^k interleaved complex doubles (host; set-up of the mask response only)
static void host_fft_pow2(std::vector<double>& a, int n) {
    for (int i = 1, j = 0; i < n; ++i) {
        int bit = n >> 1;
        for (; j & bit; bit >>= 1) j ^= bit;
        j ^= bit;
        if (i < j) { std::swap(a[2 * i], a[2 * j]); std::swap(a[2 * i + 1], a[2 * j + 1]); }
    }
    for (int len = 2; len <= n; len <<= 1) {
        const double ang = -6.283185307179586476925286766559 / len;
        for (int i = 0; i < n; i += len)
            for (int k = 0; k < len / 2; ++k) {
                const double wr = std::cos(ang * k), wi = std::sin(ang * k);
                const int u = i + k, v = i + k + len / 2;
                const double xr = a[2 * v] * wr - a[2 * v + 1] * wi, xi = a[2 * v] * wi + a[2 * v + 1] * wr;
                a[2 * v] = a[2 * u] - xr; a[2 * v + 1] = a[2 * u + 1] - xi;
                a[2 * u] += xr; a[2 * u + 1] += xi;
            }
    }
}

// Channel-mask variant (include/wofdm.h): frames in batches -- the masked Tx streams of a batch are written to HBM (the dense
// tensor-core product of mask_gemm.cu, or tx_mask_kernel with WOFDM_MASK_FFT=1), a K1 kernel reads them
// (BerParams::tx_stream) and does channel, noise, Rx and counting as always.
int wofdm_ber_run_masked(wofdm_handle h, const wofdm_sys_t* sys, const double* win_tx, const double* win_rx,
                         const double* chan, int L, int C, const double* snr_db, int n_snr, int64_t ensemble,
                         uint64_t seed, uint32_t variant, int roll_off,
                         int64_t* bit_err, int64_t* bit_tot, int64_t* sym_err, int64_t* sym_tot) {
    NvtxRange nvtx_("wofdm_ber_run_masked");
    if (!h) return WOFDM_EINVAL;
    int rc = validate_sys(h, sys, L);
    if (rc) return rc;
    if (!win_tx || !win_rx || !chan || !snr_db || !bit_err || !bit_tot || !sym_err || !sym_tot) return fail(h, WOFDM_EINVAL, "NULL buffer");
    if (C < 1 || n_snr < 1 || ensemble < 1) return fail(h, WOFDM_EINVAL, "C, n_snr and ensemble must be >= 1");
    if (sys->precision != 0) return fail(h, WOFDM_EUNSUPPORTED, "the channel-mask variant runs in fp32");
    const int N = sys->N, n_tx = N + sys->cp + sys->cs, stride = n_tx - sys->tail_tx, M = 2 * n_tx - 1;
    if (N != 128 && N != 256 && N != 512) return fail(h, WOFDM_EUNSUPPORTED, "the channel-mask variant is built for N = 128, 256, 512");
    if (roll_off < 1 || M / 2 + 2 * roll_off > M) return fail(h, WOFDM_EINVAL, "roll_off does not fit the mask");
    if (2 * n_tx > 3 * N + 1) return fail(h, WOFDM_EUNSUPPORTED, "cp + cs too long for the mask kernel (n_tx <= 1.5 N)");
    DeviceCtx& d0 = h->devs[0];              // (variant selection: the devices of a handle are alike)
    WOFDM_CUDA(h, cudaSetDevice(d0.dev));
    Choice ch, prod;
    // Tx side: the dense tensor-core product of mask_gemm.cu; WOFDM_MASK_FFT=1 keeps the per-symbol FFT kernel (mask_kernel.cuh)
    const char* mask_env = getenv("WOFDM_MASK_FFT");
    const bool use_gemm = !(mask_env && atoi(mask_env) == 1);
    const char* dump_env = getenv("WOFDM_MASK_DUMP");
    // a second-generation tensor-core kernel that gathers its stream from the product's output itself (no assembled copy in
    // HBM) where the frame fits one, else a tx_stream kernel (tuned or staged)
    bool fused = false;
    if (use_gemm && !(dump_env && *dump_env)) {
        rc = choose_variant(h, *sys, L, false, false, d0.smem_optin, &ch, nullptr, true, true, false, 1, true);
        fused = rc == WOFDM_OK && ch.var->txy;
    }
    if (!fused) rc = choose_variant(h, *sys, L, false, false, d0.smem_optin, &ch, nullptr, true, true);
    if (rc) return rc;
    if (ch.var->CL > 1) return fail(h, WOFDM_EUNSUPPORTED, "no channel-mask variant for cluster kernels");
    rc = choose_variant(h, *sys, L, false, false, d0.smem_optin, &prod, win_tx);     // noise numbering of wofdm_ber_run / _draws
    if (rc) return rc;
    if (prod.var->CL > 1) return fail(h, WOFDM_EUNSUPPORTED, "no channel-mask variant for cluster kernels");
    int nb = 0;
    long long max_ctas = 0;
    rc = prepare_kernel(h, *ch.var, ch.lay.bytes, d0.sm_count, &nb, &max_ctas);
    if (rc) return rc;
    std::vector<double> gd;
    if (!use_gemm) {
        // g = IDFT_M(ifftshift(windowRC)), main_channel_mask.m:404-412, 477-493
        std::vector<double> wrc(M, 0.0);
        {
            const int wl = M / 2, rest = M - wl - 2 * roll_off, zl = rest / 2;
            for (int i = 0; i < roll_off; ++i) {
                const double ax = -(roll_off + 1) / 2.0 + 1.0 + i;
                const double v = std::sin(1.5707963267948966 * (0.5 + ax / roll_off));
                wrc[zl + i] = v * v;
                wrc[zl + roll_off + wl + (roll_off - 1 - i)] = v * v;
            }
            for (int i = 0; i < wl; ++i) wrc[zl + roll_off + i] = 1.0;
        }
        gd.resize(2 * (size_t)M);
        std::vector<double> phc(M), phs(M);
        for (int i = 0; i < M; ++i) {
            const double a = 6.283185307179586476925286766559 * (double)i / (double)M;
            phc[i] = std::cos(a); phs[i] = std::sin(a);
        }
        for (int n = 0; n < M; ++n) {
            double re = 0.0, im = 0.0;
            int idx = 0;                                         // k n mod M
            for (int k = 0; k < M; ++k) {                        // ifftshift: shifted[k] = wrc[(k + M/2) mod M] (odd M: floor)
                const double w = wrc[(k + M / 2) % M];
                if (w != 0.0) { re += w * phc[idx]; im += w * phs[idx]; }
                idx += n;
                if (idx >= M) idx -= M;
            }
            gd[2 * n] = re / M; gd[2 * n + 1] = im / M;
        }
    }
    // the circular convolution (mod M) of an n_tx-sample symbol = a linear one with the periodic extension of g on
    // [-(n_tx-1), M-1]: Gp = FFT_P of that sequence laid out circularly in P = 8N >= 4 n_tx - 3 points, times 1/P
    const int P = 8 * N;
    std::vector<float> g(2 * (size_t)P, 0.f);
    std::vector<unsigned char> twp;
    if (!use_gemm) {
        std::vector<double> gc(2 * (size_t)P, 0.0);
        for (int mm = -(n_tx - 1); mm <= M - 1; ++mm) {
            const int gi = ((mm % M) + M) % M, ci = ((mm % P) + P) % P;
            gc[2 * ci] = gd[2 * gi]; gc[2 * ci + 1] = gd[2 * gi + 1];
        }
        host_fft_pow2(gc, P);
        for (size_t i = 0; i < g.size(); ++i) g[i] = (float)(gc[i] / P);
        cast_any(false, build_twiddles(P), twp);
    } else {
        twp.resize(16);
    }
    HostTables t;
    build_tables(*sys, win_tx, win_rx, t, ch.var->ntile > 0);   // (tx_mask_kernel shares the table: a common factor of its stream)
    std::vector<unsigned char> hchan, hsnr;
    cast_chan(false, chan, (size_t)L * C, hchan, L, ch.var->ntile > 0);
    std::vector<double> lin(n_snr);
    for (int i = 0; i < n_snr; ++i) lin[i] = std::pow(10.0, -0.1 * snr_db[i]);
    cast_any(false, lin, hsnr);
    const long long total = (long long)n_snr * C * ensemble;
    const size_t body = (size_t)sys->tail_tx + (size_t)sys->S * stride;
    const long long batch = std::min<long long>(total, use_gemm ? 16384 : 8192);
    const size_t scratch_elems = (size_t)ch.lay.pad + body + 64;
    const int grid_ber = (int)std::min<long long>(batch, max_ctas);
    const size_t scratch_bytes = ch.use_global ? (size_t)grid_ber * 2 * scratch_elems * sizeof(float2) : 0;
    // every device of the handle takes a contiguous range of the frames (frame ids are global: same draws, same counters as
    // on one device); its work is enqueued on its own stream, the counters meet on the host
    const int nd = (int)std::min<long long>((long long)h->devs.size(), std::max<long long>(1, total / 64));
    std::vector<std::vector<unsigned long long>> cnts(nd, std::vector<unsigned long long>((size_t)n_snr * 2, 0ull));
    std::vector<void*> d_cnts(nd, nullptr);
    auto run_dev = [&](DeviceCtx& d, long long f_lo, long long f_hi, void** d_cnt_out) -> int {
    WOFDM_CUDA(h, cudaSetDevice(d.dev));
    if (&d != &d0) {
        int nb_ = 0; long long mc_ = 0;
        const int rcp = prepare_kernel(h, *ch.var, ch.lay.bytes, d.sm_count, &nb_, &mc_);
        if (rcp) return rcp;
    }
    MaskGemm mg;
    const size_t mg_bytes = use_gemm ? mask_gemm_plan(*sys, batch, mg) : 0;
    int rc = arena_reserve(h, d, t.wtx.size() + t.wrx.size() + t.tw.size() + twp.size() + hchan.size() + hsnr.size() + g.size() * 4 +
                                 (size_t)n_snr * 16 + (fused ? 16 : (size_t)batch * body * sizeof(float2)) + scratch_bytes + mg_bytes + 8192);
    if (rc) return rc;
    auto put = [&](const void* src, size_t bytes, void** dst) -> cudaError_t {
        *dst = arena_take(d, std::max<size_t>(bytes, 16));
        if (!*dst) return cudaErrorMemoryAllocation;
        return src ? cudaMemcpyAsync(*dst, src, bytes, cudaMemcpyHostToDevice, d.stream) : cudaSuccess;
    };
    void *d_wtx, *d_wrx, *d_tw, *d_twp, *d_chan, *d_snr, *d_g, *d_cnt, *d_stream, *d_scr = nullptr;
    WOFDM_CUDA(h, put(t.wtx.data(), t.wtx.size(), &d_wtx));
    WOFDM_CUDA(h, put(t.wrx.data(), t.wrx.size(), &d_wrx));
    WOFDM_CUDA(h, put(t.tw.data(), t.tw.size(), &d_tw));
    WOFDM_CUDA(h, put(hchan.data(), hchan.size(), &d_chan));
    WOFDM_CUDA(h, put(hsnr.data(), hsnr.size(), &d_snr));
    WOFDM_CUDA(h, put(g.data(), g.size() * 4, &d_g));
    WOFDM_CUDA(h, put(twp.data(), twp.size(), &d_twp));
    WOFDM_CUDA(h, put(nullptr, (size_t)n_snr * 16, &d_cnt));
    WOFDM_CUDA(h, put(nullptr, fused ? 16 : (size_t)batch * body * sizeof(float2), &d_stream));
    if (scratch_bytes) WOFDM_CUDA(h, put(nullptr, scratch_bytes, &d_scr));
    WOFDM_CUDA(h, cudaMemsetAsync(d_cnt, 0, (size_t)n_snr * 16, d.stream));
    if (use_gemm) {
        rc = mask_gemm_setup(h, d, mg, roll_off, static_cast<const float*>(d_wtx));
        if (rc) return rc;
    }

    MaskParams mp;
    memset(&mp, 0, sizeof(mp));
    mp.N = N; mp.cp = sys->cp; mp.cs = sys->cs; mp.tail_tx = sys->tail_tx; mp.bits = sys->bits; mp.S = sys->S;
    mp.n_tx = n_tx; mp.stride = stride; mp.constellation = sys->constellation; mp.guard = sys->guard; mp.M = M;
    mp.win_tx = static_cast<const float*>(d_wtx); mp.tw = static_cast<const float2*>(d_tw);
    mp.Gp = static_cast<const float2*>(d_g); mp.twp = static_cast<const float2*>(d_twp);
    mp.seed = seed; mp.stream = static_cast<float2*>(d_stream);
    BerParams prm;
    fill_sys(prm, *sys, L);
    prm.chunk = ch.var->TC > 0 ? ch.chunk : prod.chunk; prm.use_global = ch.use_global;
    fill_flat(prm, *sys, win_tx, win_rx);
    prm.flat_tx = 0;                       // (the Tx stream comes from tx_mask_kernel)
    fill_split(prm, *ch.var);
    if (ch.var->TC == 0) {                 // staged kernel: the noise numbering of wofdm_ber_run / _draws (noise_at follows it)
        BerParams tmp = prm;
        fill_split(tmp, *prod.var);
        prm.n48 = tmp.n48;
    }
    prm.win_tx = d_wtx; prm.win_rx = d_wrx; prm.tw = d_tw; prm.chan = d_chan; prm.snr_lin = d_snr;
    prm.C = C; prm.n_snr = n_snr; prm.ensemble = ensemble; prm.seed = seed; prm.variant = variant;
    philox_round_keys(seed, prm.rk);
    prm.counters = static_cast<unsigned long long*>(d_cnt);
    prm.scratch = d_scr; prm.scratch_elems = (long long)scratch_elems;
    prm.tx_stream = static_cast<const float2*>(d_stream);
    if (fused) { prm.tx_y = mg.Y; prm.tx_yp = mg.Yp; }
    size_t msm = 0;
    for (long long f0 = f_lo; f0 < f_hi; f0 += batch) {
        const long long nf = std::min(batch, f_hi - f0);
        mp.frame_begin = f0; mp.frame_step = 1; mp.n_frames = nf;
        const int mgrid = (int)std::min<long long>(nf, (long long)d.sm_count);
        cudaError_t e = cudaSuccess;
        if (use_gemm) {
            rc = mask_gemm_batch(h, d, mg, seed, f0, nf, fused ? nullptr : static_cast<float2*>(d_stream));
            if (rc) return rc;
        } else
        switch (N) {
            case 128: msm = MaskSmem<128>::bytes(sys->S, stride, sys->tail_tx);
                      e = cudaFuncSetAttribute(tx_mask_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)msm);
                      if (e == cudaSuccess) tx_mask_kernel<128><<<mgrid, 256, msm, d.stream>>>(mp); break;
            case 256: msm = MaskSmem<256>::bytes(sys->S, stride, sys->tail_tx);
                      e = cudaFuncSetAttribute(tx_mask_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)msm);
                      if (e == cudaSuccess) tx_mask_kernel<256><<<mgrid, 256, msm, d.stream>>>(mp); break;
            default:  msm = MaskSmem<512>::bytes(sys->S, stride, sys->tail_tx);
                      e = cudaFuncSetAttribute(tx_mask_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)msm);
                      if (e == cudaSuccess) tx_mask_kernel<512><<<mgrid, 256, msm, d.stream>>>(mp); break;
        }
        WOFDM_CUDA(h, e);
        WOFDM_CUDA(h, cudaGetLastError());
        if (f0 == 0) {
            // development aid: WOFDM_MASK_DUMP=<file> writes the masked Tx streams of the first (up to 4) frames as raw float32 pairs
            const char* dump = dump_env;
            if (dump && *dump) {
                std::vector<float> hs(2 * body * (size_t)std::min<long long>(nf, 4));
                WOFDM_CUDA(h, cudaMemcpyAsync(hs.data(), d_stream, hs.size() * 4, cudaMemcpyDeviceToHost, d.stream));
                WOFDM_CUDA(h, cudaStreamSynchronize(d.stream));
                if (FILE* fp = fopen(dump, "wb")) { fwrite(hs.data(), 4, hs.size(), fp); fclose(fp); }
            }
        }
        prm.frame_begin = f0; prm.frame_step = 1; prm.n_frames = nf;
        WOFDM_CUDA(h, ch.var->launch(prm, (int)std::min<long long>(nf, max_ctas), ch.lay.bytes, d.stream));
        h->launches += 2;
    }
    *d_cnt_out = d_cnt;
    return WOFDM_OK;
    };   // run_dev
    for (int i = 0; i < nd; ++i) {
        rc = run_dev(h->devs[i], total * i / nd, total * (i + 1) / nd, &d_cnts[i]);
        if (rc) break;
    }
    // (the downloads into pageable memory block the host: only after every device has its work)
    for (int i = 0; i < nd; ++i) {           // (also after a failure: nothing of this call stays in flight)
        cudaSetDevice(h->devs[i].dev);
        cudaError_t es = cudaSuccess;
        if (!rc && d_cnts[i]) es = cudaMemcpyAsync(cnts[i].data(), d_cnts[i], (size_t)n_snr * 16, cudaMemcpyDeviceToHost, h->devs[i].stream);
        if (es == cudaSuccess) es = cudaStreamSynchronize(h->devs[i].stream);
        if (es != cudaSuccess && !rc) rc = fail(h, WOFDM_ECUDA, std::string("wofdm_ber_run_masked: ") + cudaGetErrorString(es));
    }
    cudaSetDevice(d0.dev);
    if (rc) return rc;
    for (int i = 0; i < n_snr; ++i) {
        unsigned long long be = 0, se = 0;
        for (int k = 0; k < nd; ++k) { be += cnts[k][2 * i]; se += cnts[k][2 * i + 1]; }
        bit_err[i] = (int64_t)be; sym_err[i] = (int64_t)se;
        sym_tot[i] = (int64_t)C * ensemble * (N - 2 * sys->guard) * (sys->S - 1);
        bit_tot[i] = sym_tot[i] * sys->bits;
    }
    return WOFDM_OK;
}

}  // extern "C"
