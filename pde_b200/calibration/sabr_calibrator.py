"""SABR smile calibrator with the reference's API (src/python/quant_trading/calibration/sabr_calibrator.py)
over the batched GPU objective (pde_b200.sabr.BatchSABR -> hb_sabr_objective).

Same formula and objective as the reference, value for value (``sabr_implied_vol`` :159-258, weighted sum of
squared vol errors :316-324, uniform weights normalised :291-293, default bounds :130-134, beta fixed).  The
optimiser is new: the reference runs SLSQP once per maturity with a finite-difference gradient in Python;
here ALL maturities are searched at once -- a scrambled-Sobol population over the box plus the reference's
own starting point (:296-306), then shrinking boxes around each smile's incumbent, one launch per round for
every (maturity, candidate).  The optimiser trajectory has no reference counterpart; losses and vols do.
"""
from __future__ import annotations

import logging
import time
from dataclasses import dataclass, field
from datetime import datetime
from typing import Dict, List, Optional, Tuple

import numpy as np

from .heston_calibrator import CalibrationError

logger = logging.getLogger(__name__)
_NAMES = ("alpha", "rho", "nu")


@dataclass
class SABRParameters:
    """sabr_calibrator.py:43-71"""

    alpha: float
    beta: float
    rho: float
    nu: float

    def __post_init__(self):
        if self.alpha <= 0:
            raise ValueError(f"alpha must be positive, got {self.alpha}")
        if not 0 <= self.beta <= 1:
            raise ValueError(f"beta must be in [0, 1], got {self.beta}")
        if not -1 < self.rho < 1:
            raise ValueError(f"rho must be in (-1, 1), got {self.rho}")
        if self.nu <= 0:
            raise ValueError(f"nu must be positive, got {self.nu}")

    def to_dict(self) -> Dict[str, float]:
        return {"alpha": self.alpha, "beta": self.beta, "rho": self.rho, "nu": self.nu}


@dataclass
class SABRCalibrationResult:
    """sabr_calibrator.py:73-104"""

    params_by_maturity: Dict[float, SABRParameters]
    rmse_by_maturity: Dict[float, float]
    total_rmse: float
    calibration_time: float
    n_maturities: int
    n_options: int
    success: bool
    message: str
    timestamp: datetime = field(default_factory=datetime.utcnow)

    def to_dict(self) -> Dict:
        return {
            "params_by_maturity": {str(T): p.to_dict() for T, p in self.params_by_maturity.items()},
            "rmse_by_maturity": {str(T): r for T, r in self.rmse_by_maturity.items()},
            "total_rmse": self.total_rmse, "calibration_time": self.calibration_time,
            "n_maturities": self.n_maturities, "n_options": self.n_options, "success": self.success,
            "message": self.message, "timestamp": self.timestamp.isoformat(),
        }


class SABRCalibrator:
    """Per-maturity SABR (alpha, rho, nu) fit, beta fixed (sabr_calibrator.py:107-157)."""

    DEFAULT_BOUNDS = {"alpha": (0.001, 2.0), "rho": (-0.99, 0.99), "nu": (0.001, 3.0)}

    def __init__(self, beta: float = 0.5, bounds: Optional[Dict[str, Tuple[float, float]]] = None, db_session=None,
                 *, device: int = 0, population: int = 8192, rounds: int = 16, round_size: int = 2048, seed: int = 42):
        if not 0 <= beta <= 1:
            raise ValueError(f"beta must be in [0, 1], got {beta}")
        self.beta = beta
        self.bounds = {**self.DEFAULT_BOUNDS, **(bounds or {})}
        self.db_session = db_session
        self._cached_params: Dict[str, Dict[float, SABRParameters]] = {}
        self.device, self.population, self.rounds, self.round_size, self.seed = device, population, rounds, round_size, seed
        self._engines: Dict[float, object] = {}

    def _engine(self, beta: float):
        from ..sabr import BatchSABR

        if beta not in self._engines:
            self._engines[beta] = BatchSABR(beta, self.device)
        return self._engines[beta]

    # ---- formula (value-for-value the reference's, evaluated on the device) -----------------------------
    def sabr_implied_vol(self, F: float, K: float, T: float, alpha: float, beta: float, rho: float, nu: float) -> float:
        return float(self._engine(beta).vols_host([[alpha, rho, nu]], [K], F, T, "py")[0, 0])

    def _sabr_atm_vol(self, F: float, T: float, alpha: float, beta: float, rho: float, nu: float) -> float:
        return self.sabr_implied_vol(F, F, T, alpha, beta, rho, nu)

    # ---- batched fit of many smiles ---------------------------------------------------------------------
    def _fit(self, strikes: List[np.ndarray], vols: List[np.ndarray], forwards, maturities,
             weights: Optional[List[np.ndarray]], guesses: List[Optional[Dict[str, float]]]):
        """-> params [n_smiles, 3], objective values [n_smiles]."""
        import torch
        from scipy.stats import qmc

        eng = self._engine(self.beta).set_smiles(strikes, vols, forwards, maturities, weights)
        dev = torch.device("cuda", self.device)
        m = len(strikes)
        lb = torch.tensor([self.bounds[k][0] for k in _NAMES], dtype=torch.float64, device=dev)
        ub = torch.tensor([self.bounds[k][1] for k in _NAMES], dtype=torch.float64, device=dev)
        sob = qmc.Sobol(d=3, seed=self.seed)
        unit = torch.as_tensor(sob.random(self.population), device=dev)  # shared by all smiles
        X = (lb + (ub - lb) * unit).unsqueeze(0).repeat(m, 1, 1)
        # the reference's starting point (:296-306) rides along as candidate 0
        for i in range(m):
            g = guesses[i]
            if g:
                x0 = [g.get("alpha", 0.3), g.get("rho", -0.3), g.get("nu", 0.5)]
            else:
                atm = int(np.argmin(np.abs(strikes[i] - forwards[i])))
                x0 = [vols[i][atm] * forwards[i] ** (1 - self.beta), -0.3, 0.5]
            X[i, 0] = torch.minimum(torch.maximum(torch.tensor(x0, dtype=torch.float64, device=dev), lb), ub)
        loss = eng.objective(X)
        loss = torch.where(torch.isfinite(loss), loss, torch.full_like(loss, float("inf")))
        best_loss, idx = loss.min(dim=1)
        best = X[torch.arange(m, device=dev), idx]
        unit_r = torch.as_tensor(sob.random(self.round_size), device=dev) * 2.0 - 1.0  # [-1, 1]^3
        half = (ub - lb) * 0.25
        for _ in range(self.rounds):
            cand = torch.minimum(torch.maximum(best.unsqueeze(1) + half * unit_r.unsqueeze(0), lb), ub)
            cand[:, 0] = best  # keep the incumbent
            loss = eng.objective(cand)
            loss = torch.where(torch.isfinite(loss), loss, torch.full_like(loss, float("inf")))
            best_loss, idx = loss.min(dim=1)
            best = cand[torch.arange(m, device=dev), idx]
            half = half * 0.4
        return best.cpu().numpy(), best_loss.cpu().numpy()

    def calibrate_single_maturity(self, strikes: np.ndarray, market_vols: np.ndarray, F: float, T: float,
                                  weights: Optional[np.ndarray] = None,
                                  initial_guess: Optional[Dict[str, float]] = None) -> Tuple[SABRParameters, float]:
        """sabr_calibrator.py:260-361: -> (parameters, RMSE)."""
        strikes = np.asarray(strikes, dtype=np.float64)
        market_vols = np.asarray(market_vols, dtype=np.float64)
        if len(strikes) < 3:
            raise CalibrationError(f"Need at least 3 strikes for SABR calibration, got {len(strikes)}")
        x, _ = self._fit([strikes], [market_vols], [F], [T], None if weights is None else [np.asarray(weights, float)],
                         [initial_guess])
        alpha, rho, nu = (float(v) for v in x[0])
        model = self._engine(self.beta).vols_host([x[0]], strikes, F, T, "py")[0]
        rmse = float(np.sqrt(np.mean((model - market_vols) ** 2)))
        return SABRParameters(alpha=alpha, beta=self.beta, rho=rho, nu=nu), rmse

    def calibrate(self, market_options, F0: float, r: float = 0.0, q: float = 0.0, use_forward: bool = True,
                  warm_start: Optional[Dict[float, Dict[str, float]]] = None,
                  underlying: Optional[str] = None) -> SABRCalibrationResult:
        """sabr_calibrator.py:363-497: all maturities in one batched search."""
        start = time.time()
        maturities = sorted(market_options["T"].unique())
        n_options = len(market_options)
        Ks, Vs, Ws, Fs, Ts, guesses, skipped = [], [], [], [], [], [], []
        fit_idx = []  # strikes each smile is FITTED on: the objective kernel stages at most 512 strikes of a smile
        for T in maturities:
            d = market_options[market_options["T"] == T]
            if len(d) < 3:  # the reference logs the CalibrationError and records rmse = inf (:470-472)
                skipped.append(T)
                continue
            Ks.append(d["strike"].values.astype(np.float64))
            Vs.append(d["implied_vol"].values.astype(np.float64))
            Ws.append(d["weight"].values.astype(np.float64) if "weight" in d.columns else np.ones(len(d)))
            # a longer smile is fitted on 512 strikes spread evenly over it (the reference has no such limit; its
            # fit quality below is still measured on every strike)
            fit_idx.append(np.arange(len(d)) if len(d) <= 512 else
                           np.unique(np.round(np.linspace(0, len(d) - 1, 512)).astype(int)))
            Fs.append(F0 * np.exp((r - q) * T) if use_forward else F0)
            Ts.append(T)
            guesses.append(warm_start.get(T) if warm_start else None)
        params_by_maturity: Dict[float, SABRParameters] = {}
        rmse_by_maturity: Dict[float, float] = {T: float("inf") for T in skipped}
        total_errors: List[float] = []
        if Ts:
            x, _ = self._fit([k[ix] for k, ix in zip(Ks, fit_idx)], [v[ix] for v, ix in zip(Vs, fit_idx)], Fs, Ts,
                             [w[ix] for w, ix in zip(Ws, fit_idx)], guesses)
            eng = self._engine(self.beta)
            for i, T in enumerate(Ts):
                model = eng.vols_host([x[i]], Ks[i], Fs[i], T, "py")[0]
                err = (model - Vs[i]) ** 2
                params_by_maturity[T] = SABRParameters(alpha=float(x[i, 0]), beta=self.beta, rho=float(x[i, 1]),
                                                       nu=float(x[i, 2]))
                rmse_by_maturity[T] = float(np.sqrt(np.mean(err)))
                total_errors.extend(err)
        success = len(params_by_maturity) == len(maturities)
        result = SABRCalibrationResult(
            params_by_maturity=params_by_maturity, rmse_by_maturity=rmse_by_maturity,
            total_rmse=float(np.sqrt(np.mean(total_errors))) if total_errors else float("inf"),
            calibration_time=time.time() - start, n_maturities=len(maturities), n_options=n_options, success=success,
            message="Calibration successful" if success else "Partial calibration")
        if underlying:
            self._cached_params[underlying] = params_by_maturity
        return result

    def get_implied_vol(self, F: float, K: float, T: float, params: Optional[SABRParameters] = None,
                        underlying: Optional[str] = None) -> float:
        """sabr_calibrator.py:499-531"""
        if params is None:
            if underlying and underlying in self._cached_params:
                cached = self._cached_params[underlying]
                params = cached[min(cached.keys(), key=lambda x: abs(x - T))]
            else:
                raise ValueError("No parameters provided and no cached params available")
        return self.sabr_implied_vol(F, K, T, params.alpha, params.beta, params.rho, params.nu)

    def interpolate_params(self, T: float, params_by_maturity: Dict[float, SABRParameters]) -> SABRParameters:
        """sabr_calibrator.py:533-582: alpha linear in total variance, rho and nu linear."""
        mats = sorted(params_by_maturity.keys())
        if T <= mats[0]:
            return params_by_maturity[mats[0]]
        if T >= mats[-1]:
            return params_by_maturity[mats[-1]]
        for i in range(len(mats) - 1):
            if mats[i] <= T <= mats[i + 1]:
                T1, T2 = mats[i], mats[i + 1]
                p1, p2 = params_by_maturity[T1], params_by_maturity[T2]
                break
        w = (T - T1) / (T2 - T1)
        var_T = p1.alpha ** 2 * T1 + w * (p2.alpha ** 2 * T2 - p1.alpha ** 2 * T1)
        return SABRParameters(alpha=float(np.sqrt(var_T / T)), beta=self.beta, rho=p1.rho + w * (p2.rho - p1.rho),
                              nu=p1.nu + w * (p2.nu - p1.nu))

    @staticmethod
    def generate_synthetic_smile(F: float = 100.0, T: float = 0.25, alpha: float = 0.3, beta: float = 0.5,
                                 rho: float = -0.3, nu: float = 0.5, n_strikes: int = 11,
                                 strike_range: Tuple[float, float] = (0.8, 1.2), noise_std: float = 0.0):
        """sabr_calibrator.py:611-659"""
        import pandas as pd

        strikes = np.linspace(F * strike_range[0], F * strike_range[1], n_strikes)
        vols = SABRCalibrator(beta=beta)._engine(beta).vols_host([[alpha, rho, nu]], strikes, F, T, "py")[0]
        if noise_std > 0:
            vols = vols + np.random.normal(0, noise_std, len(vols))
            vols = np.maximum(vols, 0.01)
        return pd.DataFrame({"strike": strikes, "T": T, "implied_vol": vols})
