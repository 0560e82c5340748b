"""w-ofdm-optimization_b200 -- B200-native hot path of felipescoelho/w-ofdm-optimization.

Only what the hot path needs: ``csrc/`` (hand-written sm_100a CUDA kernels + the C-ABI,
built into ``libwofdm.so``), ``capi`` (ctypes binding of include/wofdm.h) and ``ofdm_utils``
(host-side mirror of the reference's ``ofdm_utils.simulation_fun`` / ``interf_power``); next rows:
``optimizers``, ``channel_model``, ``timefreq``.

The directory name contains hyphens, so import it as ``import wofdm_b200`` (alias module at the
repo root) or ``importlib.import_module("w-ofdm-optimization_b200")``.
"""
from . import capi, sharding  # noqa: F401
from . import ofdm_utils  # noqa: F401,E402
from . import channel_model  # noqa: F401,E402
from . import optimizers  # noqa: F401,E402
from . import timefreq  # noqa: F401,E402
from .capi import Handle, BerPlan, SysT, WofdmError, params_from_name  # noqa: F401

__all__ = ["capi", "sharding", "ofdm_utils", "channel_model", "optimizers", "timefreq", "Handle", "BerPlan", "SysT", "WofdmError", "params_from_name"]
