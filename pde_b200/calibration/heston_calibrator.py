"""Heston calibrator with the reference's API, objective evaluated on a B200.

Counterpart of src/python/quant_trading/calibration/heston_calibrator.py: same public
classes (``HestonParameters``, ``CalibrationResult``, ``CalibrationError``,
``HestonCalibrator``), constructor arguments, ``calibrate`` signature, result fields, data
validation and failure semantics.  What changes is how the objective is evaluated:

* reference: one Python->C++ call per option per candidate, candidates strictly sequential
  (``updating="immediate"``, ``workers=1``; heston_calibrator.py:416-426, :572-584);
* here: a whole Differential-Evolution generation is priced in one GPU launch
  (``vectorized=True, updating="deferred"``), and every Levenberg-Marquardt iterate gets its
  residuals and SciPy-rule 2-point Jacobian from one launch.

Per-candidate values (prices, loss, residuals, Jacobian) match the reference; optimiser
trajectories are not a parity target (SciPy version and update order differ).

``mode="refgrid"`` (default) prices with the reference's own arithmetic (1023-point
quadrature, heston.cpp:94-151); ``mode="fft"`` uses the Carr-Madan N-point FFT the reference
documents but never implemented.
"""
from __future__ import annotations

import logging
import time
from dataclasses import dataclass, field
from datetime import datetime
from typing import TYPE_CHECKING, Any, Dict, List, Optional, Tuple

import numpy as np

if TYPE_CHECKING:  # pragma: no cover
    import pandas as pd

logger = logging.getLogger(__name__)

_NAMES = ("kappa", "theta", "sigma", "rho", "v0")


class CalibrationError(Exception):
    """Raised when model calibration fails (heston_calibrator.py:40-43)."""


@dataclass
class HestonParameters:
    """Validated Heston parameter set (heston_calibrator.py:46-128)."""

    kappa: float
    theta: float
    sigma: float
    rho: float
    v0: float

    def __post_init__(self):
        for name in ("kappa", "theta", "sigma", "v0"):
            if getattr(self, name) <= 0:
                raise ValueError(f"{name} must be positive")
        if not -1 < self.rho < 1:
            raise ValueError("rho must be in (-1, 1)")

    @property
    def feller_condition_value(self) -> float:
        return 2 * self.kappa * self.theta - self.sigma ** 2

    @property
    def is_feller_satisfied(self) -> bool:
        return self.feller_condition_value >= 0

    @property
    def feller_condition_satisfied(self) -> bool:
        return self.is_feller_satisfied

    def to_dict(self) -> Dict[str, float]:
        d = {k: getattr(self, k) for k in _NAMES}
        d["feller_satisfied"] = self.feller_condition_satisfied
        return d

    def to_array(self) -> np.ndarray:
        return np.array([getattr(self, k) for k in _NAMES])

    @classmethod
    def from_array(cls, arr) -> "HestonParameters":
        return cls(*(float(arr[i]) for i in range(5)))

    @classmethod
    def from_dict(cls, d: Dict[str, float]) -> "HestonParameters":
        return cls(*(d[k] for k in _NAMES))


@dataclass
class CalibrationResult:
    """heston_calibrator.py:131-173."""

    params: HestonParameters
    fit_quality: Dict[str, float]
    convergence: Dict[str, Any]
    timestamp: datetime
    warnings: List[str] = field(default_factory=list)

    @property
    def success(self) -> bool:
        return self.convergence.get("local_converged", False) or self.convergence.get("cached", False)

    @property
    def rmse(self) -> float:
        return self.fit_quality.get("rmse", float("inf"))

    def to_dict(self) -> Dict[str, Any]:
        return {"params": self.params.to_dict(), "fit_quality": self.fit_quality, "convergence": self.convergence,
                "timestamp": self.timestamp, "warnings": self.warnings, "success": self.success, "rmse": self.rmse}


class HestonCalibrator:
    """Two-stage (Differential Evolution -> least squares) Heston calibrator on a B200."""

    DEFAULT_BOUNDS = {  # heston_calibrator.py:201-207
        "kappa": (0.1, 10.0),
        "theta": (0.01, 1.0),
        "sigma": (0.01, 2.0),
        "rho": (-0.99, 0.99),
        "v0": (0.01, 1.0),
    }

    def __init__(self, db=None, bounds: Optional[Dict[str, Tuple[float, float]]] = None, global_maxiter: int = 100,
                 global_popsize: int = 15, local_method: str = "trf", local_ftol: float = 1e-8, *,
                 mode: str = "refgrid", n_grid: int = 4096, eta: float = 0.25, alpha: float = 0.75, device: int = 0):
        self.db = db
        self.bounds = bounds or self.DEFAULT_BOUNDS.copy()
        self.global_maxiter = global_maxiter
        self.global_popsize = global_popsize
        self.local_method = local_method
        self.local_ftol = local_ftol
        self.mode, self.n_grid, self.eta, self.alpha, self.device = mode, n_grid, eta, alpha, device
        self._pricer = None
        self._surface_key = None

    # ---- GPU plumbing -----------------------------------------------------------------------------
    def _bounds_arrays(self):
        lb = np.array([self.bounds[k][0] for k in _NAMES], dtype=np.float64)
        ub = np.array([self.bounds[k][1] for k in _NAMES], dtype=np.float64)
        return lb, ub

    def _bind_surface(self, strikes, maturities, market_prices, is_calls, S0, r, q):
        """(Re)upload the option surface when it changes; returns the BatchPricer."""
        from ..pricer import BatchPricer

        strikes = np.ascontiguousarray(strikes, dtype=np.float64)
        maturities = np.ascontiguousarray(np.broadcast_to(np.asarray(maturities, dtype=np.float64), strikes.shape))
        calls = np.ascontiguousarray(np.broadcast_to(np.asarray(is_calls), strikes.shape).astype(np.uint8))
        mkt = None if market_prices is None else np.ascontiguousarray(market_prices, dtype=np.float64)
        key = (strikes.tobytes(), maturities.tobytes(), calls.tobytes(), None if mkt is None else mkt.tobytes(),
               float(S0), float(r), float(q), tuple(self._bounds_arrays()[0]), tuple(self._bounds_arrays()[1]))
        if self._pricer is None:
            self._pricer = BatchPricer(self.mode, self.n_grid, self.eta, self.alpha, self.device)
        if key != self._surface_key:
            self._pricer.set_surface(strikes, maturities, calls, mkt, S0=S0, r=r, q=q)
            self._pricer.set_bounds(*self._bounds_arrays())
            self._surface_key = key
        return self._pricer

    # ---- batched entry points (new, additive) --------------------------------------------------------
    def price_surface_batch(self, X, strikes, maturities, is_calls, S0, r, q):
        """X [P,5] -> prices [P,n].  NumPy in -> NumPy out; CUDA tensor in -> CUDA tensor out."""
        pr = self._bind_surface(strikes, maturities, None, is_calls, S0, r, q)
        return pr.price_host(X) if isinstance(X, np.ndarray) else pr.price(X)

    def objective_batch(self, X, strikes, maturities, market_prices, is_calls, S0, r, q):
        """X [P,5] -> loss [P] with the reference's 1e10 rule; invalid sets give 1e10 instead of raising."""
        pr = self._bind_surface(strikes, maturities, market_prices, is_calls, S0, r, q)
        return pr.objective_host(X) if isinstance(X, np.ndarray) else pr.objective(X)

    def residuals_batch(self, X, strikes, maturities, market_prices, is_calls, S0, r, q):
        """X [P,5] -> (residuals [P,n], jacobian [P,n,5]) with SciPy's bounded 2-point rule."""
        pr = self._bind_surface(strikes, maturities, market_prices, is_calls, S0, r, q)
        return pr.jacobian_host(X) if isinstance(X, np.ndarray) else pr.jacobian(X)

    def normal_equations_batch(self, X, strikes, maturities, market_prices, is_calls, S0, r, q):
        """X [P,5] -> [P,22] = loss, ||r||^2, J^T r, triu(J^T J)."""
        pr = self._bind_surface(strikes, maturities, market_prices, is_calls, S0, r, q)
        return pr.normal_equations_host(X) if isinstance(X, np.ndarray) else pr.normal_equations(X)

    # ---- the reference's per-candidate methods --------------------------------------------------------
    def _price_options(self, params_array, strikes, maturities, is_calls, S0, r, q) -> np.ndarray:
        """heston_calibrator.py:538-586.  Invalid parameters raise ValueError like the reference's
        wrapper constructor (models/heston.py:166); invalid options price to NaN (:583-584)."""
        x = np.asarray(params_array, dtype=np.float64)
        from ..models.heston import HestonParameters as _P

        _P(*x[:5]).validate()
        ic = is_calls if hasattr(is_calls, "__getitem__") else np.full(len(strikes), bool(is_calls))
        pr = self._bind_surface(strikes, maturities, None, ic, S0, r, q)
        return pr.price_host(x[None, :])[0]

    def _compute_objective(self, params_array, strikes, maturities, market_prices, is_calls, S0, r, q) -> float:
        """heston_calibrator.py:486-513."""
        model_prices = self._price_options(params_array, strikes, maturities, is_calls, S0, r, q)
        if np.any(np.isnan(model_prices)) or np.any(model_prices <= 0):
            return 1e10
        errors = (model_prices - market_prices) / market_prices
        return np.sum(errors ** 2)

    def _compute_residuals(self, params_array, strikes, maturities, market_prices, is_calls, S0, r, q) -> np.ndarray:
        """heston_calibrator.py:515-536."""
        model_prices = self._price_options(params_array, strikes, maturities, is_calls, S0, r, q)
        model_prices = np.maximum(model_prices, 1e-10)
        return (model_prices - market_prices) / market_prices

    # ---- calibrate ---------------------------------------------------------------------------------
    def calibrate(self, market_options: "pd.DataFrame", S0: float, r: float, q: float,
                  warm_start: Optional[Dict[str, float]] = None, use_cached_on_failure: bool = True,
                  underlying: Optional[str] = None) -> CalibrationResult:
        """Same contract as heston_calibrator.py:247-370."""
        logger.info(f"Starting Heston calibration with {len(market_options)} options")
        t0 = time.time()
        self._validate_market_data(market_options)
        if underlying is None:
            underlying = market_options["underlying"].iloc[0] if "underlying" in market_options.columns else "UNKNOWN"
        try:
            strikes = market_options["strike"].values.astype(np.float64)
            maturities = market_options["maturity"].values.astype(np.float64)
            market_prices = market_options["mid_price"].values.astype(np.float64)
            if "is_call" in market_options.columns:
                is_calls = market_options["is_call"].values
            elif "option_type" in market_options.columns:
                is_calls = (market_options["option_type"].str.lower() == "call").values
            else:
                is_calls = np.ones(len(market_options), dtype=bool)

            logger.info("Stage 1: Global search with Differential Evolution")
            g = self._global_optimization(strikes, maturities, market_prices, is_calls, S0, r, q, warm_start)
            logger.info("Stage 2: Local refinement with Levenberg-Marquardt")
            loc = self._local_optimization(strikes, maturities, market_prices, is_calls, S0, r, q, g.x)

            params = HestonParameters.from_array(loc.x)
            warns = self._validate_parameters(params)
            for w in warns:
                logger.warning(w)
            fit = self._compute_fit_quality(params, strikes, maturities, market_prices, is_calls, S0, r, q)
            ms = int((time.time() - t0) * 1000)
            result = CalibrationResult(
                params=params, fit_quality=fit,
                convergence={"global_converged": g.success, "local_converged": loc.success, "global_nit": g.nit,
                             "local_nfev": loc.nfev, "calibration_time_ms": ms},
                timestamp=datetime.now(), warnings=warns)
            if self.db:
                self._store_results(result, underlying)
            logger.info(f"Calibration successful: RMSE={fit['rmse']:.4f}, R²={fit['r_squared']:.4f}, Time={ms}ms")
            return result
        except Exception as e:
            logger.error(f"Calibration failed: {e}")
            if use_cached_on_failure and self.db:
                logger.warning("Attempting to use cached parameters from database")
                cached = self._load_cached_parameters(underlying)
                if cached:
                    return cached
            raise CalibrationError(f"Calibration failed: {e}") from e

    def _global_optimization(self, strikes, maturities, market_prices, is_calls, S0, r, q, warm_start):
        """Differential Evolution, one GPU launch per generation (reference: :372-433)."""
        from scipy.optimize import differential_evolution

        pricer = self._bind_surface(strikes, maturities, market_prices, is_calls, S0, r, q)

        def objective(x):  # vectorized: x is (5, S)
            X = np.ascontiguousarray(np.atleast_2d(x.T))
            out = pricer.objective_host(X)
            return out if x.ndim > 1 else float(out[0])

        bounds_list = [self.bounds[k] for k in _NAMES]
        x0 = HestonParameters.from_dict(warm_start).to_array() if warm_start else None
        res = differential_evolution(objective, bounds=bounds_list, maxiter=self.global_maxiter,
                                     popsize=self.global_popsize, seed=42, x0=x0, updating="deferred",
                                     vectorized=True, polish=False)
        logger.debug(f"Global optimization: converged={res.success}, obj={res.fun:.6f}, nit={res.nit}")
        return res

    def _local_optimization(self, strikes, maturities, market_prices, is_calls, S0, r, q, x0):
        """Bounded least squares; residuals and the SciPy-rule Jacobian come from the GPU
        (reference: :435-484, SciPy builds the same 2-point Jacobian from 5 extra evaluations)."""
        from scipy.optimize import least_squares

        pricer = self._bind_surface(strikes, maturities, market_prices, is_calls, S0, r, q)
        lower, upper = self._bounds_arrays()
        mkt = np.asarray(market_prices, dtype=np.float64)

        def residuals(x):
            p = np.maximum(pricer.price_host(x[None, :])[0], 1e-10)
            return (p - mkt) / mkt

        def jac(x):
            return pricer.jacobian_host(x[None, :])[1][0]

        res = least_squares(residuals, x0=x0, jac=jac, bounds=(lower, upper), method=self.local_method,
                            ftol=self.local_ftol, xtol=1e-8, verbose=0)
        logger.debug(f"Local optimization: converged={res.success}, cost={res.cost:.6f}, nfev={res.nfev}")
        return res

    # ---- fit metrics / validation / persistence: host logic of the reference ------------------------------
    def _compute_fit_quality(self, params: HestonParameters, strikes, maturities, market_prices, is_calls, S0, r,
                             q) -> Dict[str, float]:
        """Same keys as heston_calibrator.py:588-643."""
        model = self._price_options(params.to_array(), strikes, maturities, is_calls, S0, r, q)
        err = model - market_prices
        rmse = float(np.sqrt(np.mean(err ** 2)))
        ss_res = np.sum(err ** 2)
        ss_tot = np.sum((market_prices - np.mean(market_prices)) ** 2)
        return {
            "rmse": rmse,
            "r_squared": float(1 - ss_res / ss_tot if ss_tot > 0 else 0),
            "relative_rmse": float(rmse / np.mean(market_prices)),
            "max_abs_error": float(np.max(np.abs(err))),
            "mean_abs_error": float(np.mean(np.abs(err))),
            "n_options": len(market_prices),
            "feller_satisfied": params.is_feller_satisfied,
            "feller_value": params.feller_condition_value,
        }

    def _validate_parameters(self, params: HestonParameters) -> List[str]:
        """Warnings of heston_calibrator.py:645-674."""
        out = []
        if not params.is_feller_satisfied:
            out.append(f"Feller condition violated: 2κθ = {2 * params.kappa * params.theta:.4f} < "
                       f"σ² = {params.sigma ** 2:.4f}. Variance may reach zero.")
        if params.kappa > 8.0:
            out.append(f"Very high mean-reversion speed: κ={params.kappa:.2f}")
        if params.sigma > 1.5:
            out.append(f"Very high vol of vol: σ={params.sigma:.2f}")
        if abs(params.rho) > 0.95:
            out.append(f"Extreme correlation: ρ={params.rho:.2f}")
        if params.v0 > 0.5:
            out.append(f"Very high initial variance: v₀={params.v0:.2f}")
        return out

    def _validate_market_data(self, market_options: "pd.DataFrame") -> None:
        """heston_calibrator.py:676-698."""
        for col in ("strike", "maturity", "mid_price"):
            if col not in market_options.columns:
                raise ValueError(f"Missing required column: {col}")
        if len(market_options) < 5:
            logger.warning(f"Very few options for calibration: {len(market_options)}. "
                           "Recommend at least 20 options for reliable calibration.")
        n_bad = int((market_options["mid_price"] <= 0).sum())
        if n_bad:
            raise ValueError(f"Found {n_bad} options with price <= 0")
        n_bad = int((market_options["maturity"] <= 0).sum())
        if n_bad:
            raise ValueError(f"Found {n_bad} options with maturity <= 0")

    def _store_results(self, result: CalibrationResult, underlying: str) -> None:
        """Duck-typed TimeSeriesDB.store_model_parameters (heston_calibrator.py:700-710; db.py:374)."""
        self.db.store_model_parameters(model_type="heston", underlying=underlying, parameters=result.params.to_dict(),
                                       fit_quality=result.fit_quality, maturity=None,
                                       converged=result.convergence["local_converged"],
                                       calibration_time_ms=result.convergence["calibration_time_ms"])

    def _load_cached_parameters(self, underlying: str) -> Optional[CalibrationResult]:
        """heston_calibrator.py:712-733."""
        cached = self.db.get_latest_model_parameters(model_type="heston", underlying=underlying, maturity=None)
        if cached and cached.get("converged", False):
            return CalibrationResult(params=HestonParameters.from_dict(cached["parameters"]),
                                     fit_quality=cached["fit_quality"], convergence={"cached": True},
                                     timestamp=cached["time"], warnings=["Using cached parameters"])
        return None

    @classmethod
    def generate_synthetic_data(cls, S0: float = 100.0, r: float = 0.05, q: float = 0.02, kappa: float = 2.0,
                                theta: float = 0.04, sigma: float = 0.3, rho: float = -0.7, v0: float = 0.04,
                                n_strikes: int = 11, n_maturities: int = 3, noise_std: float = 0.0,
                                strikes: Optional[np.ndarray] = None, maturities: Optional[np.ndarray] = None,
                                mode: str = "refgrid", device: int = 0) -> "pd.DataFrame":
        """Synthetic call surface with the reference's conventions (heston_calibrator.py:735-813):
        maturity-major order, optional relative N(0, noise_std) noise with a 0.01 floor, NumPy's
        global RNG consumed once per option in that order."""
        import pandas as pd

        from ..pricer import BatchPricer

        if strikes is None:
            strikes = np.linspace(0.8 * S0, 1.2 * S0, n_strikes)
        if maturities is None:
            maturities = np.linspace(0.1, 1.0, n_maturities)
        K = np.tile(np.asarray(strikes, dtype=np.float64), len(maturities))
        T = np.repeat(np.asarray(maturities, dtype=np.float64), len(strikes))
        pricer = BatchPricer(mode, device=device).set_surface(K, T, True, None, S0=S0, r=r, q=q)
        prices = pricer.price_host(np.array([[kappa, theta, sigma, rho, v0]]))[0]
        pricer.close()
        rows = []
        for Ki, Ti, price in zip(K, T, prices):
            price = float(price)
            if noise_std > 0:
                price *= 1 + np.random.normal(0, noise_std)
                price = max(price, 0.01)
            rows.append({"strike": Ki, "maturity": Ti, "mid_price": price, "option_type": "call",
                         "underlying": "SYNTHETIC", "is_call": True})
        return pd.DataFrame(rows)

    generate_synthetic_options = generate_synthetic_data
