"""Host-side mirror of the reference's time-frequency / out-of-band-radiation analysis (next row 8f-3), backed by
libwofdm.so.  Same names, tuple, file outputs and dictionary keys as

* ``ofdm_utils.timefreq_fun(data)``                      python/ofdm_utils/timefreq_simulation.py:17-76
* ``wOFDMSystem(...).estimate_obr(win_tx_mat, Ts)``      python/ofdm_utils/timefreq_simulation.py:216-296
* ``wOFDMSystem(...).analytical_psd(...)``               python/ofdm_utils/timefreq_simulation.py:150-214

The PSD estimates (X_est_*) come from the device (wofdm_psd_estimate: the record's Tx streams and their 8N-point
periodograms, `records` independent records averaged; the reference evaluates one); the analytical curves S_* are a
few closed-form vectors evaluated here with numpy/scipy as the reference does.  Documented deviation: the symbols are
the on-device Philox draws (the same ones for the optimised, RC and CP signals, like the reference's shared X).
"""
from __future__ import annotations

import os

import numpy as np

from . import capi
from .ofdm_utils import default_handle, _diag, TX_SYSTEMS

GUARD_BAND = 48      # estimate_obr's constants (:219-221)
NO_SYMBOLS = 256


class wOFDMSystem:
    def __init__(self, system_design: str, dft_len: int, cp_len: int, tail_tx: int, tail_rx: int, folder_path: str,
                 seed: int = 0, records: int = 1, handle=None):
        self.name = system_design
        self.dft_len, self.cp_len, self.tail_tx, self.tail_rx = dft_len, cp_len, tail_tx, tail_rx
        self.folder_path = folder_path
        self.seed, self.records, self._h = seed, records, handle
        self.cs_len = capi.params_from_name(system_design, dft_len, cp_len, tail_tx, tail_rx).cs   # :136-149

    def _rc_window(self):
        s = capi.params_from_name(self.name, self.dft_len, self.cp_len, self.tail_tx, self.tail_rx)
        n_tx = self.dft_len + self.cp_len + self.cs_len
        # gen_rc_window_tx(N, cp, cs, tail_tx) (transmitter.py:61-87): RC tails of tail_tx samples whatever the system
        if self.tail_tx == 0:
            return np.ones(n_tx)
        if s.tail_tx == self.tail_tx:
            return capi.rc_window_tx(s)
        a = np.arange(self.tail_tx) - (self.tail_tx - 1) / 2.0
        t = np.sin(np.pi / 2 * (0.5 + a / self.tail_tx)) ** 2
        return np.concatenate([t, np.ones(n_tx - 2 * self.tail_tx), t[::-1]])

    def analytical_psd(self, sampling_period: float, win_tx_mat, guard_band: int, fft_len: int):
        """S_opt, S_rc, S_cp (:150-214): |G|^2 of the band-pass interpolation filter times the window, scaled, times the
        correlation of the window with its own cyclic extension."""
        from scipy.signal import firwin
        N, cp, cs = self.dft_len, self.cp_len, self.cs_len
        up = fft_len / N
        f_axis = np.linspace(-.5, .5 - 1 / fft_len, fft_len) / sampling_period
        delta_f = 1 / (N * sampling_period)
        sigma2 = (N / (N - guard_band)) ** 2
        n_tx = N + cp + cs
        g = firwin(n_tx, [f_axis[int(fft_len / 2 + up)], f_axis[-int(guard_band * up)]], window="boxcar",
                   fs=1 / sampling_period, pass_zero=False)
        ripple = np.cos(f_axis / delta_f)

        def curve(w):
            G = np.abs(np.fft.fftshift(np.fft.fft(g * w, fft_len))) ** 2
            scale = G * (N * sigma2 / n_tx) / up
            c_cp = np.sum(w[:cp] * w[N:N + cp])
            c_cs = np.sum(w[N + cp:N + cp + cs] * w[cp:cp + cs])
            return scale * (np.sum(w ** 2) + 2 * (c_cp + c_cs) * ripple)

        G_cp = np.abs(np.fft.fftshift(np.fft.fft(g, fft_len))) ** 2
        S_cp = G_cp * (N * sigma2 / (N + cp)) / up * ((N + cp) + 2 * cp * ripple)
        return curve(_diag(win_tx_mat)), curve(self._rc_window()), S_cp

    def estimate_obr(self, win_tx_mat, samp_period: float):
        """(opt, rc, cp) dictionaries with the reference's keys (:283-296)."""
        h = self._h or default_handle()
        N, cp, cs = self.dft_len, self.cp_len, self.cs_len
        fft_len = 8 * N
        tail = self.tail_tx if self.name in TX_SYSTEMS else 0           # :241-246: overlap-add only for the Tx-windowed systems
        kw = dict(guard_band=GUARD_BAND, n_sym=NO_SYMBOLS, records=self.records, seed=self.seed)
        X = {"opt": h.psd_estimate(N, cp, cs, tail, _diag(win_tx_mat), **kw),
             "rc": h.psd_estimate(N, cp, cs, tail, self._rc_window(), **kw),
             "cp": h.psd_estimate(N, cp, cs, 0, np.ones(N + cp + cs), **kw)}
        f_axis = np.linspace(-.5, .5 - (1 / fft_len), fft_len) / (200 * 1e-9)
        interp = fft_len / N
        gb = int(interp * GUARD_BAND)
        S = dict(zip(("opt", "rc", "cp"), self.analytical_psd(samp_period, win_tx_mat, GUARD_BAND, fft_len)))
        out = []
        for k in ("opt", "rc", "cp"):
            x = X[k]
            obr = np.mean(np.hstack((x[:gb], x[-gb:])))
            mf = np.hstack((x[gb:int(fft_len / 2)], x[-int(fft_len / 2 - interp):-gb]))
            out.append({f"X_est_{k}": x, f"S_{k}": S[k], "f_axis": f_axis, f"obr_{k}": obr, f"mf_band_{k}": mf})
        return tuple(out)


def timefreq_fun(data: tuple, seed: int = 0, records: int = 1, handle=None):
    """Drop-in for ofdm_utils.timefreq_fun: same 7-tuple (system_design, dft_len, cp_len, tail_tx, tail_rx, window_path,
    folder_path), same window files, same timefreq/{opt,rc}_<sys>_<cp>.npz and CP_<cp>.npz outputs."""
    system_design, dft_len, cp_len, tail_tx, tail_rx = data[0:5]
    window_path, folder_path = data[5:]
    model = wOFDMSystem(system_design, dft_len, cp_len, tail_tx, tail_rx, folder_path, seed=seed, records=records, handle=handle)
    n_tx = dft_len + cp_len + model.cs_len
    if system_design in TX_SYSTEMS:                                     # :47-54: the Tx tail of the stored reduced variable
        x = np.load(os.path.join(window_path, f"{system_design}_{cp_len}.npy")).ravel()[:tail_tx + 1]
        win_tx = np.concatenate([x[:0:-1], np.full(n_tx - 2 * tail_tx, x[0]), x[1:]])   # reduce_variable_tx (utils.py:13-43)
    elif system_design in ("wrx", "CPwrx", "CP"):
        win_tx = np.ones(n_tx)
    else:
        raise ValueError(f"unknown w-OFDM system {system_design!r}")
    opt, rc, cpd = model.estimate_obr(win_tx, 200 * 1e-9)
    path = os.path.join(folder_path, "timefreq")
    os.makedirs(path, exist_ok=True)
    np.savez(os.path.join(path, f"opt_{system_design}_{cp_len}.npz"), **opt)
    np.savez(os.path.join(path, f"rc_{system_design}_{cp_len}.npz"), **rc)
    np.savez(os.path.join(path, f"CP_{cp_len}.npz"), **cpd)
    return opt, rc, cpd
