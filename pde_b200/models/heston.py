"""Pythonic Heston wrapper with the reference's API (src/python/quant_trading/models/heston.py).

Same class names, constructor keywords, defaults, return types and error behaviour as the
reference wrapper (models/heston.py:32-365); the C++ object it forwards to is replaced by the
B200 module :mod:`pde_b200.cpp.quant_cpp`.  There is no CPU pricing fallback (the reference has
none either, models/heston.py:159-163).
"""
from __future__ import annotations

import warnings
from dataclasses import dataclass
from typing import Dict, List, Optional, Union

import numpy as np

from ..cpp import quant_cpp


@dataclass
class HestonParameters:
    """kappa, theta, sigma, rho, v0 (models/heston.py:32-95)."""

    kappa: float
    theta: float
    sigma: float
    rho: float
    v0: float

    def is_feller_satisfied(self) -> bool:
        return 2.0 * self.kappa * self.theta >= self.sigma ** 2

    def is_valid(self) -> bool:
        return self.kappa > 0 and self.theta > 0 and self.sigma > 0 and abs(self.rho) < 1 and self.v0 > 0

    def validate(self) -> None:
        checks = (("kappa", self.kappa > 0, "kappa must be positive"), ("theta", self.theta > 0, "theta must be positive"),
                  ("sigma", self.sigma > 0, "sigma must be positive"), ("rho", abs(self.rho) < 1, "|rho| must be < 1"),
                  ("v0", self.v0 > 0, "v0 must be positive"))
        for name, ok, text in checks:
            if not ok:
                raise ValueError(f"{text}, got {getattr(self, name)}")

    def to_dict(self) -> Dict[str, float]:
        return {k: getattr(self, k) for k in ("kappa", "theta", "sigma", "rho", "v0")}


@dataclass
class OptionGreeks:
    delta: float
    gamma: float
    vega: float
    theta: float
    rho: float


@dataclass
class PricingResult:
    price: float
    greeks: Optional[OptionGreeks] = None


class HestonModel:
    """Heston pricer facade (models/heston.py:118-371)."""

    def __init__(self, kappa: float = 2.0, theta: float = 0.04, sigma: float = 0.3, rho: float = -0.7,
                 v0: float = 0.04):
        self.params = HestonParameters(kappa, theta, sigma, rho, v0)
        self.params.validate()
        if not self.params.is_feller_satisfied():
            # same warning class and condition as models/heston.py:168-173
            warnings.warn(f"Feller condition violated: 2κθ = {2 * kappa * theta:.4f}, σ² = {sigma ** 2:.4f}. "
                          "Variance may hit zero.", UserWarning)
        self._cpp_model = quant_cpp.heston.HestonModel(quant_cpp.heston.HestonParameters(kappa, theta, sigma, rho, v0))

    def price_option(self, strike: float, maturity: float, spot: float, rate: float, dividend: float = 0.0,
                     is_call: bool = True) -> float:
        return self._cpp_model.price_option(strike, maturity, spot, rate, dividend, is_call)

    def price_option_with_greeks(self, strike: float, maturity: float, spot: float, rate: float,
                                 dividend: float = 0.0, is_call: bool = True) -> PricingResult:
        r = self._cpp_model.price_option_with_greeks(strike, maturity, spot, rate, dividend, is_call)
        g = r.greeks
        return PricingResult(price=r.price, greeks=OptionGreeks(g.delta, g.gamma, g.vega, g.theta, g.rho))

    def price_options(self, strikes: Union[List[float], np.ndarray], maturities: Union[List[float], np.ndarray, float],
                      spot: float, rate: float, dividend: float = 0.0, is_call: bool = True) -> np.ndarray:
        mats = [float(maturities)] if isinstance(maturities, (int, float)) else list(maturities)
        return np.array(self._cpp_model.price_options(list(strikes), mats, spot, rate, dividend, is_call))

    def implied_volatility(self, strike: float, maturity: float, spot: float, rate: float, dividend: float = 0.0,
                           is_call: bool = True) -> float:
        return self._cpp_model.implied_volatility(strike, maturity, spot, rate, dividend, is_call)

    def implied_volatility_surface(self, strikes, maturities, spot: float, rate: float,
                                   dividend: float = 0.0) -> np.ndarray:
        """(strikes x maturities) matrix of call implied vols (models/heston.py:313-343)."""
        # the reference loops over implied_volatility per (K, T); one batched launch gives the same numbers
        from ..pricer import BatchPricer

        strikes, maturities = np.asarray(strikes, dtype=np.float64), np.asarray(maturities, dtype=np.float64)
        K = np.repeat(strikes, len(maturities))
        T = np.tile(maturities, len(strikes))
        p = self.params
        pricer = BatchPricer("refgrid", device=quant_cpp._DEVICE).set_surface(K, T, True, None, S0=spot, r=rate,
                                                                              q=dividend)
        iv = pricer.implied_vol_host(np.array([[p.kappa, p.theta, p.sigma, p.rho, p.v0]]))[0]
        pricer.close()
        return iv.reshape(len(strikes), len(maturities))

    @classmethod
    def from_dict(cls, params: Dict[str, float]) -> "HestonModel":
        return cls(**{k: params[k] for k in ("kappa", "theta", "sigma", "rho", "v0")})

    @classmethod
    def from_params(cls, params: HestonParameters) -> "HestonModel":
        return cls(params.kappa, params.theta, params.sigma, params.rho, params.v0)

    def __repr__(self) -> str:
        p = self.params
        return f"HestonModel(κ={p.kappa:.3f}, θ={p.theta:.4f}, σ={p.sigma:.3f}, ρ={p.rho:.3f}, v0={p.v0:.4f})"
