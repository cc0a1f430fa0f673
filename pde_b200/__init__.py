"""pde_b200 -- B200-native (sm_100a, FP64) batched Heston Carr-Madan pricing and calibration
objective: a drop-in for that one hot path of dharvpat/PDE.

Layout (only what the path needs):
  csrc/         hand-written CUDA kernels + the C ABI (include/heston_b200.h)
  pricer.py     BatchPricer: torch-tensor / NumPy host API over the C ABI
  cpp/          quant_cpp-compatible module object (reference: src/cpp/bindings)
  models/       HestonModel wrapper with the reference's API (reference: models/heston.py)
  calibration/  HestonCalibrator with the reference's API + batched drivers
  sabr.py       BatchSABR: batched SABR vols / smile objective
  sharding.py   parameter-set sharding across GPUs + loss all-gather
"""
from .sabr import BatchSABR  # noqa: F401
from .pricer import BatchPricer, characteristic_function, fft_batch, launch_count, measure_fp64_peak  # noqa: F401

__version__ = "0.1.0"
