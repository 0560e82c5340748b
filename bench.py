#!/usr/bin/env python
"""bench.py -- w-OFDM BER Monte-Carlo throughput (OFDM symbols/s) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # our arm (CUDA, libwofdm.so)
    python bench.py --impl reference [--steps K] [--warmup W]      # CPU arm: port of the reference's loop
    torchrun --nproc-per-node N ... bench.py --gpus N ...          # N > 1: one rank per GPU

A step = one pass of the BER chain over one batch: BASELINE.json configs[1] -- wtx-OFDM, N=256, cp=16,
tail_tx=8, 16-QAM, optimised-Tx-window stand-in, the full channel set (250 Vehicular-A realisations,
21 taps), 30 SNR points linspace(-20,50,30), 1e8 bits per point (ensemble 27 -> 202 500 frames =
3.24 M OFDM symbols per GPU per step; weak scaling: the ensemble grows with N, frames are sharded by
global frame id, one all-reduce of the int64 counters per step).
Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for every field.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

METRIC = "w-OFDM BER Monte Carlo OFDM symbols/s"
UNIT = "OFDM symbols/s"
CFG = dict(system="wtx", N=256, cp=16, tail_tx=8, tail_rx=0, bits=4, S=16, L=21, C=250, n_snr=30,
           snr_lo=-20.0, snr_hi=50.0, ensemble_per_gpu=27, noise_norm=1, constellation=1,
           name="configs[1]: wtx-OFDM N=256 cp=16 tail_tx=8 16-QAM (Gray, BER), optimised-Tx-window stand-in, "
                "250 VehA channels x 30 SNR points x ensemble 27 per GPU = 202500 frames = 3.24e6 OFDM symbols "
                "(1.0e8 bits per SNR point) per GPU per step, S=16, L=21")
# --workload configs4: BASELINE.json configs[4], the scaled stress case (not the headline line): N=1024, 64-QAM, WOLA with
# 4x scaled tails, 10 000 synthetic VehA channels generated ON THE DEVICE (wofdm_gen_channels), 1e10 bits per SNR point
CFG4 = dict(system="WOLA", N=1024, cp=64, tail_tx=32, tail_rx=40, bits=6, S=16, L=21, C=10000, n_snr=30,
            snr_lo=-20.0, snr_hi=50.0, ensemble_per_gpu=11, noise_norm=1, constellation=1,
            name="configs[4]: WOLA-OFDM N=1024 cp=64 tail_tx=32 tail_rx=40 64-QAM (Gray, BER), optimised-window stand-ins, "
                 "10000 VehA channels (GMEDS_1, generated on the device) x 30 SNR points x ensemble 11 per GPU = 3.3e6 frames "
                 "= 5.28e7 OFDM symbols (1.0e10 bits per SNR point) per GPU per step, S=16, L=21")


def workload_inputs(handle=None):
    """Synthetic inputs of the named shapes (nothing ships with the reference).  Pure numpy, except the
    stress workload's 10 000 channels, which come from the device generator (K4) when a handle is given."""
    rng = np.random.default_rng(1)
    L, C = CFG["L"], CFG["C"]
    if handle is not None and C > 1000:
        fd, fs, frame = (100 / 3.6 / 299792458.0) * 2e9, 5e6, 16 * 256 * 200e-9        # wofdm_optimization.py:63-76
        chan = handle.gen_channels("vehicularA", L, fd, fs, frame, no_frames=1, n_sets=C, seed=1)
        return chan, np.linspace(CFG["snr_lo"], CFG["snr_hi"], CFG["n_snr"])
    d = np.array([0.0, 310, 710, 1090, 1730, 2510]) / 200.0            # ITU-R VehA delays in samples (Ts = 200 ns)
    pw = 10.0 ** (np.array([0.0, -1, -9, -10, -15, -20]) / 10.0)
    g = (rng.standard_normal((6, C)) + 1j * rng.standard_normal((6, C))) * np.sqrt(pw / 2)[:, None]
    axis = np.arange(L) - (L - 1) / 2.0
    chan = np.sinc(d[None, :] - axis[:, None]) @ g                    # (L, C) complex128
    snr = np.linspace(CFG["snr_lo"], CFG["snr_hi"], CFG["n_snr"])
    return chan, snr


def flops_per_symbol(N, n_tx, stride, tail_rx, L):
    """algorithmic real flops per OFDM symbol."""
    f_fft = 2 * 5 * N * int(math.log2(N))
    f_conv = 8 * L * stride
    f_misc = 2 * n_tx + 2 * (N + tail_rx) + 2 * tail_rx + 4 * stride + 8 * stride + 8 * N
    return f_fft, f_conv, f_misc


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thr = threading.Thread(target=self._read, daemon=True)
            self.thr.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.25)
            self.proc.terminate()
            self.thr.join(timeout=2)

    def summary(self):
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        busy = sorted(sm)[len(sm) // 2:]          # upper half = samples taken under load
        return {"sm_mhz": float(np.median(busy)), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def cpu_port_throughput(budget_s, workers):
    """The CPU port of the reference's loop (oracle/wofdm_cpu_port.py) on a bounded sample of the workload."""
    from oracle import wofdm_cpu_port as P
    chan, snr = workload_inputs()
    c = CFG

    def tasks_for(ens, n_tasks, seed0):
        return [P.build_task(c["system"], c["N"], c["cp"], c["tail_tx"], c["tail_rx"], c["S"], c["bits"],
                             chan[:, (k % c["C"]):(k % c["C"]) + 1], ens, snr, seed0 + k) for k in range(n_tasks)]
    # calibrate on one core: 30 SNR points x 1 channel x ensemble 1 = 480 symbols
    rate1, _, _, _ = P.timed_throughput(tasks_for(1, 1, 0), 1)
    n_tasks = max(workers, 1) * 2
    ens = max(1, int(rate1 * budget_s * 0.8 / (c["n_snr"] * c["S"] * 2)))
    val, dt, syms, _ = P.timed_throughput(tasks_for(ens, n_tasks, 100), workers)
    sample = (f"{n_tasks} tasks x ({c['n_snr']} SNR x 1 channel x ensemble {ens}) = {syms} OFDM symbols of the "
              f"configs[1] workload in {dt:.1f} s, numba dense-matrix port of wofdm_simulation.py:171-240")
    return val, sample, rate1


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    workers = os.cpu_count() or 1
    vals = []
    sample = ""
    for i in range(args.warmup + args.steps):
        budget = 2.0 if i < args.warmup else max(3.0, min(12.0, 120.0 / max(args.steps, 1)))
        v, sample, _ = cpu_port_throughput(budget, workers)
        if i >= args.warmup:
            vals.append(v)
    value = float(np.mean(vals))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": None, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": "configs[1]: wtx-OFDM N=256 cp=16 tail_tx=8 16-QAM, 250 VehA channels, 30 SNR points, S=16",
                       "note": "CPU arm: bounded sample per step, all host cores, one process per core"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": workers, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


_JSON_FD = None


def emit(line):
    """The one JSON line of the contract, on the process's original stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        sys.stdout.flush()
        os.write(_JSON_FD, data)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="configs1", choices=["configs1", "configs4"],
                    help="configs1 = BASELINE's headline configuration (default); configs4 = the N=1024 stress case")
    args = ap.parse_args()
    if args.workload == "configs4":
        CFG.clear(); CFG.update(CFG4)
        args.no_cpu_baseline = True          # the CPU port is dimensioned for the headline workload
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)

    # stdout carries ONE JSON line: NCCL writes its log lines to stdout by default ("NCCL version ..." was seen on the
    # multi-GPU boxes, where NCCL_DEBUG=VERSION ignores NCCL_DEBUG_FILE), so every other writer of file descriptor 1 --
    # C libraries included -- is sent to stderr and the JSON line goes out through the saved descriptor
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)
    import torch
    import torch.distributed as dist
    import wofdm_b200 as W
    from wofdm_b200 import capi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} needs WORLD_SIZE={args.gpus} (launch with torch.distributed.run); got {world}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    c = CFG
    s = W.params_from_name(c["system"], c["N"], c["cp"], c["tail_tx"], c["tail_rx"], bits=c["bits"], S=c["S"],
                           noise_norm=c["noise_norm"], constellation=c["constellation"], precision=0)
    rng = np.random.default_rng(7)
    x_tx = np.concatenate([[1.0], np.clip(capi.rc_window_tx(s)[-c["tail_tx"]:] * (1 + 0.1 * rng.uniform(-1, 1, c["tail_tx"])), 0, 1)])
    win_tx = capi.expand_window_tx(s, x_tx)                 # "optimised" Tx window stand-in
    win_rx = capi.rc_window_rx(s)
    ens_total = c["ensemble_per_gpu"] * world
    shard = (rank, world)
    frames_rank = c["n_snr"] * c["C"] * c["ensemble_per_gpu"]
    syms_rank = frames_rank * c["S"]

    h = W.Handle([local])
    chan, snr = workload_inputs(h)
    stream = torch.cuda.Stream()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")      # > 126 MB L2
    peak_tflops, peak_mhz = h.fp32_peak(0)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident arm: inputs live in HBM, one kernel launch per step ----------------
    plan = h.ber_plan(s, win_tx, win_rx, chan, snr)
    totals = np.zeros((c["n_snr"], 2), dtype=np.int64)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    with torch.cuda.stream(stream):
        for i in range(args.warmup):
            plan.launch(ens_total, seed=1000 + i, shard=shard, stream=stream.cuda_stream)
        barrier()
        launches0 = h.launches
        with ClockSampler(local) as clocks:
            for i in range(args.steps):
                flush.fill_(i & 0xff)                                       # evict L2 between timed iterations
                ev[i][0].record(stream)
                kev[i][0].record(stream)
                dptr = plan.launch(ens_total, seed=2000 + i, shard=shard, stream=stream.cuda_stream)
                kev[i][1].record(stream)
                if world > 1:                                               # the path's only exchange step
                    t = _as_tensor(torch, dptr, c["n_snr"] * 2)
                    dist.all_reduce(t)
                ev[i][1].record(stream)
            barrier()
        launches = h.launches - launches0
    ms_rank = sum(a.elapsed_time(b) for a, b in ev)
    kms = sum(a.elapsed_time(b) for a, b in kev) / args.steps
    ms_t = torch.tensor([ms_rank], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms_t, op=dist.ReduceOp.MAX)
    ms_total = float(ms_t.item())
    value = syms_rank * world * args.steps / (ms_total * 1e-3)
    be, se = plan.read()
    bt, st_ = plan.totals(ens_total, shard if world == 1 else (0, 1))
    # (N > 1: the in-place all-reduce left the job-wide counters in every rank's device buffer)

    # ---------------- end-to-end arm: C-ABI call with HOST buffers, copies inside the timed region ----------------
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
    win_tx_p, win_rx_p, chan_p, snr_p = pin(win_tx), pin(win_rx), pin(chan), pin(snr)
    for i in range(2):
        h.ber_run(s, win_tx_p, win_rx_p, chan_p, snr_p, ens_total, seed=3000 + i, shard=shard)
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        r = h.ber_run(s, win_tx_p, win_rx_p, chan_p, snr_p, ens_total, seed=4000 + i, shard=shard)
        if world > 1:
            t = torch.from_numpy(np.stack([r["bit_err"], r["sym_err"]])).cuda()
            dist.all_reduce(t)
            t.cpu()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    e2e_t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_val = syms_rank * world * args.steps / float(e2e_t.item())
    h2d = (s.n_tx + s.N + s.tail_rx) * 8 + chan.size * 16 + snr.size * 8
    d2h = c["n_snr"] * 2 * 8

    if rank == 0:
        f_fft, f_conv, f_misc = flops_per_symbol(s.N, s.n_tx, s.stride, s.tail_rx, c["L"])
        f_chain = f_fft + f_conv + f_misc
        achieved = f_chain * syms_rank / (kms * 1e-3) / 1e12
        traffic = None
        tpath = os.path.join(REPO, "profiles", "k1_traffic.json")
        if os.path.exists(tpath):
            with open(tpath) as fh:
                traffic = json.load(fh).get("dram_bytes_per_launch")
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": c["name"],
                       "kernel": plan.kernel, "frames_per_gpu_per_step": frames_rank,
                       "l2": "256 MiB buffer rewritten between timed iterations (inputs are ~100 KB; the kernel is FP32-bound)",
                       "parallelism": f"frames sharded by global id over {world} GPU(s), one int64 all-reduce per step"},
            "clocks": clocks.summary(),
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "api": "wofdm_ber_run_shard (ctypes, pinned host buffers in, int64 counters out)"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "fp32", "achieved": achieved, "peak": peak_tflops, "unit": "TFLOP/s", "frac": achieved / peak_tflops,
                         "traffic": traffic,
                         "peak_source": f"FMA-only micro-benchmark run in this process (wofdm_diag_fp32_peak, scalar FFMA): {peak_tflops:.1f} TFLOP/s "
                                        f"= 128 FMA lanes x 148 SMs at {peak_mhz:.0f} MHz; "
                                        "MEASURED_PEAKS.json has no FP32 entry (HBM/bf16 only); nominal 74.4 TFLOP/s",
                         "flops_per_symbol": {"fft": f_fft, "conv": f_conv, "misc": f_misc, "chain": f_chain},
                         "kernel_ms": kms, "achieved_fft_only": f_fft * syms_rank / (kms * 1e-3) / 1e12,
                         "hbm_note": "HBM traffic ~0 (counters only): 'hbm'/'tensor' do not bound this kernel",
                         "bound_note": "on-chip bound: the denominator is the FP32 FMA peak the north star names; the kernel's real "
                                       "limits are the math dispatch port (FFMA2/IMAD/LOP3 = 2 cycles per warp instruction, no "
                                       "co-issue of integer work with packed FP32) and shared-memory wavefronts -- DESIGN.md section 3, "
                                       "tools/ubench/pipes*.cu",
                         "tensor_note": ("kernel ber_f32t_*: the 'conv' share of the algorithmic flops (the L-tap channel convolution) "
                                         "runs on the tensor pipe (tcgen05 kind::f16, 3-term fp16 split = fp32-grade, DESIGN.md section 3) "
                                         "to relieve the FP32 dispatch port; FFTs, windows, noise, equaliser and slicer stay on the FP32 pipe; "
                                         "the figure is still algorithmic flops / time against the FP32 FMA peak"
                                         if "f32t" in plan.kernel else "all arithmetic on the FP32 pipe")},
            "ber_check": {"snr_db": [float(snr[k]) for k in (0, 10, 15, 20, 29)],
                          "ber": [float(be[k] / max(bt[k], 1)) if world == 1 else float(be[k] / (bt[k])) for k in (0, 10, 15, 20, 29)]},
        }
        if world == 1 and not args.no_cpu_baseline:
            v, sample, rate1 = cpu_port_throughput(12.0, os.cpu_count() or 1)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port", "sample": sample,
                                    "single_core": rate1}
        emit(line)
    plan.close()
    h.close()
    if world > 1:
        dist.destroy_process_group()


def _as_tensor(torch, dptr, n):
    """int64 view of the library's device counters (no copy) for the NCCL all-reduce."""
    class _Arr:
        __cuda_array_interface__ = {"shape": (n,), "typestr": "<i8", "data": (int(dptr), False), "version": 3}
    return torch.as_tensor(_Arr(), device="cuda")


if __name__ == "__main__":
    main()
