"""longphase-s_b200 — B200-native (sm_100a) read-to-variant hot path of LongPhase-S.

The directory name is not a Python identifier; load it with `__graft_entry__.load_package()`
(module name `longphase_s_b200`).  The compute path lives in csrc/ (CUDA + C ABI, built into
liblps_b200.so); this package is the thin host-side mirror used by tests and bench.py.
"""
from . import _ffi  # noqa: F401

__all__ = ["_ffi", "synth", "host"]
