"""ov_plane_b200 — B200-native MSCKF / point-on-plane EKF update hot path of rpng/ov_plane.

Layout (hot path only):
  csrc/      hand-written sm_100a CUDA kernels + the extern "C" ABI of include/ovp.h  -> lib/libovp.so
  api.py     ctypes binding of the C ABI (plumbing; no compute, no fallback)
  synth.py   deterministic synthetic clone-window scenarios (BASELINE.json configs)
  jpl.py     JPL quaternion helpers for the generator
The CUDA library is loaded lazily by `ov_plane_b200.api.lib()` and fails loudly when it has not been built.
"""
__version__ = "0.1.0"
