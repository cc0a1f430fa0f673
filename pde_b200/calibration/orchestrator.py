"""Calibration orchestrator with the reference's API (src/python/quant_trading/calibration/orchestrator.py)
over the GPU calibrators: the caller on the other side of the hot path.

Kept: ``CalibrationConfig``, ``CalibrationRunResult``, ``CalibrationStatus``, ``run_daily_calibration`` with its
status logic (:166-278), the Heston option filter (:404-447), the SABR maturity screen (:322-368), warm-start
caches, quality warnings (:449-491), ``get_cached_params`` / ``clear_cache``.
Changed on purpose:
  * the reference hands its ``T``-column option chain straight to ``HestonCalibrator.calibrate``, which requires
    ``maturity`` and ``mid_price`` (orchestrator.py:418 vs heston_calibrator.py:678) -- Heston can therefore never
    run there; here the frame is adapted (``T`` -> ``maturity``; rows without a positive ``mid_price`` dropped);
  * ``run_batch`` calibrates many underlyings in one call (the batch axis the GPU wants);
  * OU fitting is out of scope: ``spreads_data`` is recorded as a warning, not fitted;
  * no database code: ``db_session`` is only passed through to the calibrators (parameter store out of scope).
"""
from __future__ import annotations

import logging
from dataclasses import dataclass, field
from datetime import date, datetime, timezone
from enum import Enum
from typing import Dict, List, Optional

import numpy as np

from .heston_calibrator import CalibrationError, HestonCalibrator
from .sabr_calibrator import SABRCalibrator

logger = logging.getLogger(__name__)


class CalibrationStatus(Enum):
    PENDING = "pending"
    RUNNING = "running"
    SUCCESS = "success"
    PARTIAL = "partial"
    FAILED = "failed"


@dataclass
class CalibrationConfig:
    """orchestrator.py:47-71"""

    heston_enabled: bool = True
    heston_max_options: int = 100
    heston_min_options: int = 10
    heston_timeout: float = 60.0
    sabr_enabled: bool = True
    sabr_beta: float = 0.5
    sabr_min_strikes: int = 5
    ou_enabled: bool = True
    ou_min_observations: int = 60
    ou_max_half_life: float = 120.0
    use_cached_on_failure: bool = True
    cache_expiry_days: int = 5
    alert_on_failure: bool = True
    rmse_alert_threshold: float = 0.05


def _utcnow() -> datetime:
    """Naive UTC timestamp, as the reference's datetime.utcnow() (orchestrator.py:88) without its deprecation."""
    return datetime.now(timezone.utc).replace(tzinfo=None)


@dataclass
class CalibrationRunResult:
    """orchestrator.py:74-110"""

    run_date: date
    status: CalibrationStatus
    underlying: str
    heston_result: Optional[Dict] = None
    sabr_result: Optional[Dict] = None
    ou_results: Optional[Dict[str, Dict]] = None
    start_time: datetime = field(default_factory=lambda: _utcnow())
    end_time: Optional[datetime] = None
    total_time: float = 0.0
    errors: List[str] = field(default_factory=list)
    warnings: List[str] = field(default_factory=list)

    def to_dict(self) -> Dict:
        return {
            "run_date": self.run_date.isoformat(), "status": self.status.value, "underlying": self.underlying,
            "heston_result": self.heston_result, "sabr_result": self.sabr_result, "ou_results": self.ou_results,
            "start_time": self.start_time.isoformat(), "end_time": self.end_time.isoformat() if self.end_time else None,
            "total_time": self.total_time, "errors": self.errors, "warnings": self.warnings,
        }


class CalibrationOrchestrator:
    def __init__(self, config: Optional[CalibrationConfig] = None, db_session=None, *, heston_mode: str = "refgrid",
                 device: int = 0):
        self.config = config or CalibrationConfig()
        self.db_session = db_session
        self.heston_calibrator = HestonCalibrator(db=db_session, mode=heston_mode, device=device)
        self.sabr_calibrator = SABRCalibrator(beta=self.config.sabr_beta, db_session=db_session, device=device)
        self._last_heston_params: Dict[str, Dict] = {}
        self._last_sabr_params: Dict[str, Dict] = {}
        self._last_ou_params: Dict[str, Dict] = {}

    # ---- one underlying (orchestrator.py:166-278) -------------------------------------------------------
    def run_daily_calibration(self, underlying: str, options_data=None, spreads_data: Optional[Dict[str, np.ndarray]] = None,
                              S0: float = 100.0, r: float = 0.05, q: float = 0.02,
                              calibration_date: Optional[date] = None) -> CalibrationRunResult:
        start = _utcnow()
        result = CalibrationRunResult(run_date=calibration_date or date.today(), status=CalibrationStatus.RUNNING,
                                      underlying=underlying, start_time=start)
        heston_ok = sabr_ok = ou_ok = True
        if self.config.heston_enabled and options_data is not None:
            try:
                result.heston_result = self._run_heston_calibration(underlying, options_data, S0, r, q)
            except Exception as e:  # as the reference: any failure is recorded, the run goes on
                logger.error(f"Heston calibration failed: {e}")
                result.errors.append(f"Heston: {str(e)}")
                heston_ok = False
        if self.config.sabr_enabled and options_data is not None:
            try:
                result.sabr_result = self._run_sabr_calibration(underlying, options_data, S0, r, q)
            except Exception as e:
                logger.error(f"SABR calibration failed: {e}")
                result.errors.append(f"SABR: {str(e)}")
                sabr_ok = False
        if self.config.ou_enabled and spreads_data:
            result.warnings.append("OU fitting is out of scope of this build: %d spread series ignored" % len(spreads_data))
        result.end_time = _utcnow()
        result.total_time = (result.end_time - start).total_seconds()
        if heston_ok and sabr_ok and ou_ok:
            result.status = CalibrationStatus.SUCCESS
        elif heston_ok or sabr_ok or ou_ok:
            result.status = CalibrationStatus.PARTIAL
        else:
            result.status = CalibrationStatus.FAILED
        self._check_calibration_quality(result)
        return result

    def run_batch(self, chains: Dict[str, Dict], calibration_date: Optional[date] = None) -> Dict[str, CalibrationRunResult]:
        """Many underlyings in one call: ``chains[symbol] = {"options_data": df, "S0": .., "r": .., "q": ..}``."""
        return {sym: self.run_daily_calibration(sym, calibration_date=calibration_date, **kw) for sym, kw in chains.items()}

    # ---- Heston (orchestrator.py:280-320) ------------------------------------------------------------------
    @staticmethod
    def _heston_frame(options_data):
        df = options_data
        if "maturity" not in df.columns and "T" in df.columns:
            df = df.rename(columns={"T": "maturity"})
        if "mid_price" not in df.columns:
            raise CalibrationError("Heston calibration needs a mid_price column")
        return df[df["mid_price"] > 0]

    def _run_heston_calibration(self, underlying, options_data, S0, r, q) -> Dict:
        # clean first (rows without a positive mid price never reach the calibrator), then cap and count: the
        # minimum applies to what is actually calibrated on
        options_data = self._heston_frame(options_data)
        if len(options_data) > self.config.heston_max_options:
            options_data = self._filter_options_for_heston(options_data, self.config.heston_max_options)
        if len(options_data) < self.config.heston_min_options:
            raise CalibrationError(f"Insufficient options: {len(options_data)} < {self.config.heston_min_options}")
        res = self.heston_calibrator.calibrate(market_options=options_data, S0=S0, r=r, q=q,
                                               warm_start=self._last_heston_params.get(underlying),
                                               use_cached_on_failure=self.config.use_cached_on_failure,
                                               underlying=underlying)
        if res.success:
            self._last_heston_params[underlying] = res.params.to_dict()
        return res.to_dict()

    # ---- SABR (orchestrator.py:322-368) ----------------------------------------------------------------------
    def _run_sabr_calibration(self, underlying, options_data, S0, r, q) -> Dict:
        counts = options_data.groupby("T").size()
        valid = [T for T, c in counts.items() if c >= self.config.sabr_min_strikes]
        if not valid:
            raise CalibrationError(f"No maturities with >= {self.config.sabr_min_strikes} strikes")
        res = self.sabr_calibrator.calibrate(market_options=options_data[options_data["T"].isin(valid)], F0=S0, r=r, q=q,
                                             warm_start=self._last_sabr_params.get(underlying), underlying=underlying)
        if res.success:
            self._last_sabr_params[underlying] = {T: p.to_dict() for T, p in res.params_by_maturity.items()}
        return res.to_dict()

    # ---- helpers (orchestrator.py:404-547) ---------------------------------------------------------------------
    def _filter_options_for_heston(self, options_data, max_options: int):
        import pandas as pd

        col = "T" if "T" in options_data.columns else "maturity"
        maturities = sorted(options_data[col].unique())
        target = [T for T in maturities if 0.08 <= T <= 0.5] or maturities[:3]
        per = max_options // len(target)
        out = []
        for T in target:
            d = options_data[options_data[col] == T].copy()
            if "moneyness" not in d.columns:
                d["moneyness"] = abs(np.log(d["strike"] / d["strike"].median()))
            out.append(d.nsmallest(min(per, len(d)), "moneyness"))
        return pd.concat(out, ignore_index=True)

    def _check_calibration_quality(self, result: CalibrationRunResult) -> None:
        thr = self.config.rmse_alert_threshold
        if result.heston_result and result.heston_result.get("rmse", 0) > thr:
            result.warnings.append(f"Heston RMSE {result.heston_result['rmse']:.4f} exceeds threshold {thr}")
        if result.sabr_result and result.sabr_result.get("total_rmse", 0) > thr:
            result.warnings.append(f"SABR RMSE {result.sabr_result['total_rmse']:.4f} exceeds threshold {thr}")

    def get_cached_params(self, underlying: str, model_type: str) -> Optional[Dict]:
        return {"heston": self._last_heston_params, "sabr": self._last_sabr_params,
                "ou": self._last_ou_params}.get(model_type, {}).get(underlying)

    def clear_cache(self, underlying: Optional[str] = None) -> None:
        for cache in (self._last_heston_params, self._last_sabr_params):
            if underlying:
                cache.pop(underlying, None)
            else:
                cache.clear()
        if underlying:
            self._last_ou_params = {k: v for k, v in self._last_ou_params.items() if underlying not in k}
        else:
            self._last_ou_params.clear()
