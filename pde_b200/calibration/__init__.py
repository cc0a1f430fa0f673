"""Counterpart of the reference's quant_trading.calibration (Heston and SABR parts)."""
from .heston_calibrator import CalibrationError, CalibrationResult, HestonCalibrator, HestonParameters  # noqa: F401
from .population import PopulationCalibrator, sobol_population  # noqa: F401,E402
from .sabr_calibrator import SABRCalibrationResult, SABRCalibrator, SABRParameters  # noqa: F401,E402
from .orchestrator import (CalibrationConfig, CalibrationOrchestrator, CalibrationRunResult,  # noqa: F401,E402
                           CalibrationStatus)
