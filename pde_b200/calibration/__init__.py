"""Counterpart of the reference's quant_trading.calibration (Heston part only)."""
from .heston_calibrator import CalibrationError, CalibrationResult, HestonCalibrator, HestonParameters  # noqa: F401
from .population import PopulationCalibrator, sobol_population  # noqa: F401,E402
