"""Population / multi-start Heston calibration on the GPU (BASELINE.json configs 3-4).

New capability on top of the drop-in path: the reference's optimisers are sequential by
construction (``updating="immediate"``, ``workers=1``; heston_calibrator.py:416-426), so its
calibration costs ~8,000 objective evaluations one after another.  Here

1. a scrambled-Sobol population over the calibrator's box (up to millions of candidates) is priced
   in one launch per rank -- candidates are sharded across ranks, the per-candidate losses are
   all-gathered (:mod:`pde_b200.sharding`);
2. the best ``n_starts`` candidates are refined together by a bounded Levenberg-Marquardt iteration
   whose J^T J / J^T r blocks come from ``hb_normal_eq`` (SciPy's 2-point step rule on the
   reference's residuals, heston_calibrator.py:515-536), one launch per iteration for all starts.

The tiny 5x5 solves use ``torch.linalg`` on the device; all pricing arithmetic is in
libheston_b200.so.  Per-candidate losses and Jacobian blocks are the parity-tested quantities; the
optimiser trajectory is new and has no reference counterpart.
"""
from __future__ import annotations

import time
from datetime import datetime
from typing import Dict, Optional, Tuple

import numpy as np

from .heston_calibrator import CalibrationResult, HestonCalibrator, HestonParameters, _NAMES

_IU = np.triu_indices(5)


def sobol_population(n: int, lb, ub, seed: int = 42, skip: int = 0) -> np.ndarray:
    """Scrambled Sobol points scaled to the box (SURVEY.md 8d synthetic parameter sets)."""
    from scipy.stats import qmc

    s = qmc.Sobol(d=5, seed=seed)
    if skip:
        s.fast_forward(skip)
    return np.asarray(lb) + (np.asarray(ub) - np.asarray(lb)) * s.random(n)


def _unpack_normal_equations(neq):
    """[S,22] -> rr [S], g = J^T r [S,5], A = J^T J [S,5,5] (torch tensors)."""
    import torch

    S = neq.shape[0]
    A = torch.zeros((S, 5, 5), dtype=neq.dtype, device=neq.device)
    iu0 = torch.as_tensor(_IU[0], device=neq.device)
    iu1 = torch.as_tensor(_IU[1], device=neq.device)
    A[:, iu0, iu1] = neq[:, 7:]
    A[:, iu1, iu0] = neq[:, 7:]
    return neq[:, 1], neq[:, 2:7], A


class PopulationCalibrator:
    """Global population search + batched multi-start LM, one process per GPU.

    Args:
        bounds: as :class:`HestonCalibrator` (default its ``DEFAULT_BOUNDS``).
        mode, n_grid, eta, alpha, device: pricer configuration (``mode="fft"`` is the headline path).
        group: optional ``torch.distributed`` process group for sharding the population.
    """

    def __init__(self, bounds: Optional[Dict[str, Tuple[float, float]]] = None, *, mode: str = "fft",
                 n_grid: int = 4096, eta: float = 0.25, alpha: float = 0.75, device: int = 0, group=None):
        self.bounds = bounds or HestonCalibrator.DEFAULT_BOUNDS.copy()
        self.mode, self.n_grid, self.eta, self.alpha, self.device, self.group = mode, n_grid, eta, alpha, device, group
        self._pricer = None

    def _lb_ub(self):
        return (np.array([self.bounds[k][0] for k in _NAMES], dtype=np.float64),
                np.array([self.bounds[k][1] for k in _NAMES], dtype=np.float64))

    def bind(self, strikes, maturities, market_prices, is_calls, S0, r, q):
        from ..pricer import BatchPricer

        if self._pricer is None:
            self._pricer = BatchPricer(self.mode, self.n_grid, self.eta, self.alpha, self.device)
        self._pricer.set_surface(strikes, maturities, is_calls, market_prices, S0=S0, r=r, q=q)
        self._pricer.set_bounds(*self._lb_ub())
        return self._pricer

    # ---- stage 1: population -------------------------------------------------------------------------
    def population_losses(self, X):
        """X: CUDA tensor [P,5], identical on every rank -> losses [P] (sharded + all-gathered)."""
        from ..sharding import sharded_map

        return sharded_map(self._pricer.objective, X, self.group)

    # ---- stage 2: batched bounded LM ---------------------------------------------------------------------
    def refine(self, X0, iters: int = 30, mu0: float = 1e-3, tol: float = 1e-10):
        """Levenberg-Marquardt on all rows of X0 [S,5] at once.  Returns (X, rr, n_iter)."""
        import torch

        pr = self._pricer
        lb, ub = (torch.as_tensor(a, device=X0.device) for a in self._lb_ub())
        X = X0.clone()
        rr, g, A = _unpack_normal_equations(pr.normal_equations(X))
        mu = torch.full_like(rr, mu0)
        eye = torch.eye(5, dtype=X.dtype, device=X.device)
        it = 0
        for it in range(1, iters + 1):
            D = torch.diag_embed(torch.clamp(torch.diagonal(A, dim1=1, dim2=2), min=1e-12))
            M = A + mu[:, None, None] * D + 1e-14 * eye
            ok = torch.isfinite(M).all(dim=(1, 2)) & torch.isfinite(g).all(dim=1)
            M = torch.where(ok[:, None, None], M, eye)
            step = torch.linalg.solve(M, -torch.where(ok[:, None], g, torch.zeros_like(g)))
            Xt = torch.minimum(torch.maximum(X + step, lb), ub)
            rr_t, g_t, A_t = _unpack_normal_equations(pr.normal_equations(Xt))
            better = ok & torch.isfinite(rr_t) & (rr_t < rr)
            rel = torch.where(better, (rr - rr_t) / torch.clamp(rr, min=1e-300), torch.zeros_like(rr))
            X = torch.where(better[:, None], Xt, X)
            g = torch.where(better[:, None], g_t, g)
            A = torch.where(better[:, None, None], A_t, A)
            rr = torch.where(better, rr_t, rr)
            mu = torch.where(better, mu * 0.3, mu * 5.0).clamp(1e-12, 1e12)
            if bool(((rel < tol) & better | (mu >= 1e12)).all()):
                break
        return X, rr, it

    # ---- the whole thing -------------------------------------------------------------------------------------
    def calibrate(self, market_options, S0: float, r: float, q: float, n_candidates: int = 65536, n_starts: int = 32,
                  lm_iters: int = 30, seed: int = 42) -> CalibrationResult:
        """Same inputs and result type as ``HestonCalibrator.calibrate`` (heston_calibrator.py:247-370)."""
        import torch

        t0 = time.time()
        helper = HestonCalibrator(bounds=self.bounds, mode=self.mode, n_grid=self.n_grid, eta=self.eta,
                                  alpha=self.alpha, device=self.device)
        helper._validate_market_data(market_options)
        strikes = market_options["strike"].values.astype(np.float64)
        maturities = market_options["maturity"].values.astype(np.float64)
        market = market_options["mid_price"].values.astype(np.float64)
        if "is_call" in market_options.columns:
            is_calls = market_options["is_call"].values
        elif "option_type" in market_options.columns:
            is_calls = (market_options["option_type"].str.lower() == "call").values
        else:
            is_calls = np.ones(len(market_options), dtype=bool)
        self.bind(strikes, maturities, market, is_calls, S0, r, q)
        dev = torch.device("cuda", self.device)
        lb, ub = self._lb_ub()
        X = torch.as_tensor(sobol_population(n_candidates, lb, ub, seed), device=dev)
        loss = self.population_losses(X)
        usable = torch.isfinite(loss) & (loss < 1e10)
        n_usable = int(usable.sum())
        order = torch.argsort(torch.where(usable, loss, torch.full_like(loss, float("inf"))))
        starts = X[order[: max(1, min(n_starts, max(n_usable, 1)))]]
        Xr, rr, n_it = self.refine(starts, iters=lm_iters)
        best = int(torch.argmin(torch.where(torch.isfinite(rr), rr, torch.full_like(rr, float("inf")))))
        x = Xr[best].cpu().numpy()
        params = HestonParameters.from_array(x)
        warns = helper._validate_parameters(params)
        prices = self._pricer.price(Xr[best: best + 1]).cpu().numpy()[0]
        err = prices - market
        rmse = float(np.sqrt(np.mean(err ** 2)))
        ss_tot = float(np.sum((market - market.mean()) ** 2))
        fit = {"rmse": rmse, "r_squared": float(1 - np.sum(err ** 2) / ss_tot if ss_tot > 0 else 0),
               "relative_rmse": float(rmse / market.mean()), "max_abs_error": float(np.max(np.abs(err))),
               "mean_abs_error": float(np.mean(np.abs(err))), "n_options": len(market),
               "feller_satisfied": params.is_feller_satisfied, "feller_value": params.feller_condition_value}
        conv = {"global_converged": n_usable > 0, "local_converged": bool(np.isfinite(float(rr[best]))),
                "global_nit": 1, "local_nfev": int(n_it * len(starts)), "calibration_time_ms": int((time.time() - t0) * 1e3),
                "n_candidates": int(n_candidates), "n_usable_candidates": n_usable, "n_starts": int(len(starts)),
                "best_population_loss": float(loss[order[0]]), "final_sum_sq_residuals": float(rr[best])}
        return CalibrationResult(params=params, fit_quality=fit, convergence=conv, timestamp=datetime.now(), warnings=warns)
