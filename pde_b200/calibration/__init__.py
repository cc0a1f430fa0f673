"""Counterpart of the reference's quant_trading.calibration (Heston part only)."""
from .heston_calibrator import CalibrationError, CalibrationResult, HestonCalibrator, HestonParameters  # noqa: F401
