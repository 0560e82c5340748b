"""Channel-generation throughput: BASELINE configs[4]'s 10 000 synthetic channels (one-frame sets, as the reference's
driver stores them) and one long fading set; wall time of the C-ABI call with the result copied to the host."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import wofdm_b200 as W

h = W.Handle([0])
fd, fs, frame = (100 / 3.6 / 299792458.0) * 2e9, 5e6, 16 * 256 * 200e-9
for n_sets, frames in ((10000, 1), (1, 100000), (250, 1)):
    h.gen_channels("vehicularA", 21, fd, fs, frame, no_frames=frames, n_sets=n_sets, seed=1)
    t0 = time.perf_counter()
    for _ in range(5):
        c = h.gen_channels("vehicularA", 21, fd, fs, frame, no_frames=frames, n_sets=n_sets, seed=1)
    dt = (time.perf_counter() - t0) / 5
    print(f"{n_sets} sets x {frames} frames: {dt*1e3:.3f} ms per call, {n_sets*frames/dt:.3g} channels/s")
