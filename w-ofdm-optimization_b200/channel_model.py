"""Host-side mirror of the reference's channel generator, backed by libwofdm.so.

``gen_chan(standard, no_samples, doppler_freq, sampling_rate, frame_duration, no_frames)`` keeps the signature and the
result layout -- (no_samples, no_frames) complex128 -- of ``channel_model.gen_chan``
(python/channel_model/itur_channels.py:33-94); ``gen_channel_set`` is the loop of ``wofdm_optimization.py -m gen_chan``
(python/wofdm_optimization.py:63-86: no_channels independent one-frame sets) as ONE device launch, writing the same
``channels/<standard>.npy`` file.  The oscillator phases come from the device's Philox streams (seed) instead of
numpy's global generator; pass ``phases`` to inject the reference's own draws.  No CPU fallback."""
from __future__ import annotations

import os

import numpy as np

from . import ofdm_utils

SPEED_OF_LIGHT = 299792458.0    # scipy.constants.speed_of_light, as used at wofdm_optimization.py:63


def gen_chan(standard: str, no_samples: int, doppler_freq: float, sampling_rate: float, frame_duration: float,
             no_frames: int, seed: int = 0, phases=None, handle=None) -> np.ndarray:
    h = handle or ofdm_utils.default_handle()
    ph = None if phases is None else np.asarray(phases, dtype=np.float64)[None]
    return h.gen_channels(standard, no_samples, doppler_freq, sampling_rate, frame_duration, no_frames=no_frames,
                          n_sets=1, seed=seed, phases=ph)


def gen_channel_set(channel_standard: str, no_channels: int, channel_data_folder: str | None = None, seed: int = 0,
                    carrier_frequency: float = 2e9, sample_period: float = 200e-9, velocity: float = 100 / 3.6,
                    no_samples: int = 21, no_symbols: int = 16, dft_length: int = 256, handle=None) -> np.ndarray:
    """(no_samples, no_channels) complex128; saved to <channel_data_folder>/<standard>.npy when a folder is given."""
    h = handle or ofdm_utils.default_handle()
    frame_duration = no_symbols * dft_length * sample_period
    doppler_freq = (velocity / SPEED_OF_LIGHT) * carrier_frequency
    chan = h.gen_channels(channel_standard, no_samples, doppler_freq, 1 / sample_period, frame_duration, no_frames=1,
                          n_sets=no_channels, seed=seed)
    if channel_data_folder is not None:
        os.makedirs(channel_data_folder, exist_ok=True)
        np.save(os.path.join(channel_data_folder, channel_standard + ".npy"), chan)
    return chan
