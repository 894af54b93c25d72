"""B200-native scan-to-map registration for LeGO-LOAM-SR (the MapOptimization hot path).

Layout:
  csrc/    hand-written sm_100a CUDA kernels + the extern "C" ABI (include/legoloam_b200.h)
  host/    C++ mirror of the reference's MapOptimization member functions over that ABI
  abi.py   ctypes binding of the C-ABI (what tests and bench.py call)
  synth.py seeded synthetic VLP-16 / HDL-64E inputs
There is no CPU fallback: abi.py raises if the CUDA library is missing.
"""
__all__ = ["abi", "synth", "build"]
