"""Drop-in for the reference's pybind11 module ``quant_trading.cpp.quant_cpp`` -- Heston part.

Mirrors the attribute tree, signatures, keyword names, defaults, return types and error
behaviour of ``PYBIND11_MODULE(quant_cpp)`` (src/cpp/bindings/quant_cpp.cpp:27-144) /
``bind_heston`` (src/cpp/bindings/heston_bindings.cpp:15-254), with the arithmetic done on a
B200 through libheston_b200.so.  Injecting this module as ``quant_trading.cpp.quant_cpp``
lets ``quant_trading.models.heston`` run unmodified.

``price_option`` / ``price_options`` / ``characteristic_function`` are the hot path and run on
the GPU in "refgrid" mode, i.e. they return what the reference returns (heston.cpp:94-167).
``price_option_with_greeks`` and ``implied_volatility`` are the reference's host-side recipes
(heston.cpp:169-218, :275-349) over those GPU prices.

``quant_cpp.sabr`` (SABRModel, SABRParameters) evaluates the reference's Hagan formula on
the device through ``hb_sabr_vols``.  OU / PDE submodules are out of scope and absent.
"""
from __future__ import annotations

import ctypes as C
import math
import types
from typing import List, Sequence

import numpy as np

from .. import _lib
from .._lib import check

__version__ = "0.1.0"  # quant_cpp.cpp:143

_dp = C.POINTER(C.c_double)
_DEVICE = 0


def set_device(device: int) -> None:
    """CUDA ordinal used by the scalar drop-in calls of this module."""
    global _DEVICE
    _DEVICE = int(device)


def _arr(v: Sequence[float]):
    a = np.ascontiguousarray(v, dtype=np.float64)
    return a, a.ctypes.data_as(_dp)


class OptionGreeks:
    """heston.hpp:114-120 / heston_bindings.cpp:17-39."""

    def __init__(self):
        self.delta = 0.0
        self.gamma = 0.0
        self.vega = 0.0
        self.theta = 0.0
        self.rho = 0.0

    def __repr__(self) -> str:
        return ("OptionGreeks(delta=%f, gamma=%f, vega=%f, theta=%f, rho=%f)"
                % (self.delta, self.gamma, self.vega, self.theta, self.rho))


class PricingResult:
    """heston.hpp:125-129 / heston_bindings.cpp:42-55."""

    def __init__(self):
        self.price = 0.0
        self.greeks = OptionGreeks()
        self.greeks_computed = False


class HestonParameters:
    """heston.hpp:42-109 / heston_bindings.cpp:58-111."""

    def __init__(self, kappa: float = 2.0, theta: float = 0.04, sigma: float = 0.3, rho: float = -0.7,
                 v0: float = 0.04):
        self.kappa = float(kappa)
        self.theta = float(theta)
        self.sigma = float(sigma)
        self.rho = float(rho)
        self.v0 = float(v0)

    def _vec(self):
        return [self.kappa, self.theta, self.sigma, self.rho, self.v0]

    def is_feller_satisfied(self) -> bool:  # heston.hpp:65-67
        return 2.0 * self.kappa * self.theta >= self.sigma * self.sigma

    def is_valid(self) -> bool:  # heston.hpp:72-74
        return self.kappa > 0.0 and self.theta > 0.0 and self.sigma > 0.0 and abs(self.rho) < 1.0 and self.v0 > 0.0

    def validate(self) -> None:  # heston.hpp:81-100 -> ValueError with the reference's message
        _, p = _arr(self._vec())
        check(_lib.load().hb_model_validate(p))

    def __repr__(self) -> str:  # HestonParameters::to_string, heston.hpp:102-108
        return ("HestonParameters(kappa=%f, theta=%f, sigma=%f, rho=%f, v0=%f, feller=%s)"
                % (self.kappa, self.theta, self.sigma, self.rho, self.v0,
                   "OK" if self.is_feller_satisfied() else "VIOLATED"))


def _norm_cdf(x: float) -> float:  # heston.cpp:16-18
    return 0.5 * (1.0 + math.erf(x / math.sqrt(2.0)))


def _norm_pdf(x: float) -> float:  # heston.cpp:21-24
    return 0.3989422804014327 * math.exp(-0.5 * x * x)


def _bs_price(spot, strike, rate, dividend, maturity, vol, is_call):  # heston.cpp:275-293
    if maturity <= 0.0:
        return max(spot - strike, 0.0) if is_call else max(strike - spot, 0.0)
    forward = spot * math.exp((rate - dividend) * maturity)
    discount = math.exp(-rate * maturity)
    vst = vol * math.sqrt(maturity)
    d1 = (math.log(forward / strike) + 0.5 * vol * vol * maturity) / vst
    d2 = d1 - vst
    if is_call:
        return spot * math.exp(-dividend * maturity) * _norm_cdf(d1) - strike * discount * _norm_cdf(d2)
    return strike * discount * _norm_cdf(-d2) - spot * math.exp(-dividend * maturity) * _norm_cdf(-d1)


def _bs_vega(spot, strike, rate, dividend, maturity, vol):  # heston.cpp:295-309
    if maturity <= 0.0 or vol <= 0.0:
        return 0.0
    forward = spot * math.exp((rate - dividend) * maturity)
    sqrt_t = math.sqrt(maturity)
    d1 = (math.log(forward / strike) + 0.5 * vol * vol * maturity) / (vol * sqrt_t)
    return spot * math.exp(-dividend * maturity) * sqrt_t * _norm_pdf(d1)


class HestonModel:
    """heston.hpp:142-285 / heston_bindings.cpp:114-253."""

    def __init__(self, params: HestonParameters):
        params.validate()  # heston.cpp:28-30
        self._params = HestonParameters(*params._vec())

    def parameters(self) -> HestonParameters:
        return self._params

    def set_parameters(self, params: HestonParameters) -> None:  # heston.cpp:32-35
        params.validate()
        self._params = HestonParameters(*params._vec())

    def characteristic_function(self, u: complex, T: float, S0: float, r: float, q: float) -> complex:
        u = complex(u)
        _, p = _arr(self._params._vec())
        out = np.empty(2)
        check(_lib.load().hb_model_cf(p, u.real, u.imag, float(T), float(S0), float(r), float(q),
                                      out.ctypes.data_as(_dp), _DEVICE))
        return complex(out[0], out[1])

    def price_options(self, strikes: List[float], maturities: List[float], spot: float, rate: float,
                      dividend: float, is_call: bool = True) -> List[float]:
        k, kp = _arr(strikes)
        t, tp = _arr(maturities)
        _, p = _arr(self._params._vec())
        out = np.empty(k.size)
        check(_lib.load().hb_model_price_options(p, k.size, kp, t.size, tp, float(spot), float(rate), float(dividend),
                                                 int(bool(is_call)), out.ctypes.data_as(_dp), _DEVICE))
        return out.tolist()

    def price_option(self, strike: float, maturity: float, spot: float, rate: float, dividend: float,
                     is_call: bool = True) -> float:
        return self.price_options([float(strike)], [float(maturity)], spot, rate, dividend, is_call)[0]

    def price_option_with_greeks(self, strike: float, maturity: float, spot: float, rate: float, dividend: float,
                                 is_call: bool = True) -> PricingResult:
        """Finite-difference Greeks with the reference's bumps (heston.cpp:169-218)."""
        res = PricingResult()
        res.price = self.price_option(strike, maturity, spot, rate, dividend, is_call)
        eps_spot, eps_rate, eps_time, eps_vol = spot * 0.001, 0.0001, 1.0 / 365.0, 0.001
        up = self.price_option(strike, maturity, spot + eps_spot, rate, dividend, is_call)
        dn = self.price_option(strike, maturity, spot - eps_spot, rate, dividend, is_call)
        res.greeks.delta = (up - dn) / (2.0 * eps_spot)
        res.greeks.gamma = (up - 2.0 * res.price + dn) / (eps_spot * eps_spot)
        r_up = self.price_option(strike, maturity, spot, rate + eps_rate, dividend, is_call)
        r_dn = self.price_option(strike, maturity, spot, rate - eps_rate, dividend, is_call)
        res.greeks.rho = (r_up - r_dn) / (2.0 * eps_rate)
        if maturity > eps_time:
            later = self.price_option(strike, maturity - eps_time, spot, rate, dividend, is_call)
            res.greeks.theta = (later - res.price) / eps_time
        else:
            res.greeks.theta = 0.0
        pv = self._params._vec()
        m_up = HestonModel(HestonParameters(pv[0], pv[1], pv[2], pv[3], pv[4] + eps_vol))
        m_dn = HestonModel(HestonParameters(pv[0], pv[1], pv[2], pv[3], pv[4] - eps_vol))
        v_up = m_up.price_option(strike, maturity, spot, rate, dividend, is_call)
        v_dn = m_dn.price_option(strike, maturity, spot, rate, dividend, is_call)
        res.greeks.vega = (v_up - v_dn) / (2.0 * eps_vol)
        res.greeks_computed = True
        return res

    def implied_volatility(self, strike: float, maturity: float, spot: float, rate: float, dividend: float,
                           is_call: bool = True) -> float:
        """Newton iteration on Black-Scholes (heston.cpp:311-349)."""
        target = self.price_option(strike, maturity, spot, rate, dividend, is_call)
        if maturity <= 0.0:
            return 0.0
        vol = math.sqrt(self._params.v0)
        for _ in range(100):
            bs = _bs_price(spot, strike, rate, dividend, maturity, vol, is_call)
            vega = _bs_vega(spot, strike, rate, dividend, maturity, vol)
            if vega < 1e-12:
                vol *= 1.5
                continue
            diff = bs - target
            if abs(diff) < 1e-8:
                return vol
            vol = max(0.001, min(5.0, vol - diff / vega))
        return vol


# attribute tree of the reference module: quant_cpp.heston.{...}
heston = types.ModuleType(__name__ + ".heston", "Heston stochastic volatility model (B200 drop-in).")
heston.OptionGreeks = OptionGreeks
heston.PricingResult = PricingResult
heston.HestonParameters = HestonParameters
heston.HestonModel = HestonModel


# ---- quant_cpp.sabr: SABRModel (src/cpp/bindings/sabr_bindings.cpp, src/cpp/models/sabr.cpp) ----------------

class SABRParameters:
    """sabr.hpp:14-62 (defaults: typical equity parameters)."""

    def __init__(self, alpha: float = 0.2, beta: float = 0.5, rho: float = -0.3, nu: float = 0.4):
        self.alpha, self.beta, self.rho, self.nu = float(alpha), float(beta), float(rho), float(nu)

    def is_valid(self) -> bool:
        return self.alpha > 0.0 and 0.0 <= self.beta <= 1.0 and abs(self.rho) < 1.0 and self.nu >= 0.0

    def validate(self) -> None:
        if not self.alpha > 0.0:
            raise ValueError("SABR: alpha must be positive, got %f" % self.alpha)
        if not 0.0 <= self.beta <= 1.0:
            raise ValueError("SABR: beta must be in [0, 1], got %f" % self.beta)
        if not abs(self.rho) < 1.0:
            raise ValueError("SABR: |rho| must be < 1, got %f" % self.rho)
        if not self.nu >= 0.0:
            raise ValueError("SABR: nu must be non-negative, got %f" % self.nu)

    def __repr__(self) -> str:
        return "SABRParameters(alpha=%f, beta=%f, rho=%f, nu=%f)" % (self.alpha, self.beta, self.rho, self.nu)


class SABRModel:
    """Drop-in for ``quant_cpp.sabr.SABRModel``: same names, keyword arguments and exceptions; the formula
    (sabr.cpp:34-192) is evaluated on the device through ``hb_sabr_vols`` (flavour HB_SABR_CPP)."""

    def __init__(self, beta: float = 0.5):
        self._set_beta(beta)

    def _set_beta(self, beta: float) -> None:  # sabr.cpp:19-32
        if beta < 0.0 or beta > 1.0:
            raise ValueError("SABR: beta must be in [0, 1], got %f" % beta)
        self._beta = float(beta)

    beta = property(lambda self: self._beta, _set_beta)

    @staticmethod
    def _check(strike, forward, maturity, alpha, rho, nu) -> None:  # the reference's throws, sabr.cpp:132-150
        if strike is not None and not strike > 0.0:
            raise ValueError("SABR: strike must be positive")
        if not forward > 0.0:
            raise ValueError("SABR: forward must be positive")
        if not alpha > 0.0:
            raise ValueError("SABR: alpha must be positive")
        if not abs(rho) < 1.0:
            raise ValueError("SABR: |rho| must be < 1")
        if not nu >= 0.0:
            raise ValueError("SABR: nu must be non-negative")
        if not maturity >= 0.0:
            raise ValueError("SABR: maturity must be non-negative")

    def _vols(self, strikes, forward, maturity, alpha, rho, nu) -> np.ndarray:
        K, kp = _arr(strikes)
        x, xp = _arr([alpha, rho, nu])
        out = np.empty(K.size)
        check(_lib.load().hb_sabr_vols_host(_lib.HB_SABR_CPP, self._beta, float(forward), float(maturity), K.size, kp, xp,
                                            1, out.ctypes.data_as(_dp)))
        return out

    def implied_volatility(self, strike: float, forward: float, maturity: float, alpha=None, rho=None, nu=None,
                           params: "SABRParameters" = None) -> float:
        if params is not None or isinstance(alpha, SABRParameters):  # overload (strike, forward, maturity, params)
            p = params if params is not None else alpha
            alpha, rho, nu = p.alpha, p.rho, p.nu
        self._check(strike, forward, maturity, alpha, rho, nu)
        return float(self._vols([strike], forward, maturity, alpha, rho, nu)[0])

    def atm_volatility(self, forward: float, maturity: float, alpha: float, rho: float, nu: float) -> float:
        """sabr.cpp:95-128.  Evaluated as the K = F branch of implied_volatility (identical arithmetic for
        maturity >= 1e-10; below that the T-correction is < 1e-10 relative and is dropped as in :152-155)."""
        self._check(None, forward, maturity, alpha, rho, nu)
        return float(self._vols([forward], forward, maturity, alpha, rho, nu)[0])

    def implied_volatilities(self, strikes: Sequence[float], forward: float, maturity: float, alpha: float, rho: float,
                             nu: float) -> List[float]:
        for k in strikes:
            self._check(k, forward, maturity, alpha, rho, nu)
        if len(strikes) == 0:
            return []
        return [float(v) for v in self._vols(strikes, forward, maturity, alpha, rho, nu)]

    def volatility_sensitivities(self, strike: float, forward: float, maturity: float, alpha: float, rho: float,
                                 nu: float):
        """(d sigma/d alpha, d sigma/d rho, d sigma/d nu): the reference's central differences, sabr.cpp:250-279."""
        ea, er, en = alpha * 0.001, 0.001, max(nu * 0.001, 0.0001)
        iv = lambda a, r, n: self.implied_volatility(strike, forward, maturity, a, r, n)  # noqa: E731
        d_alpha = (iv(alpha + ea, rho, nu) - iv(alpha - ea, rho, nu)) / (2.0 * ea)
        r_up, r_dn = min(rho + er, 0.999), max(rho - er, -0.999)
        d_rho = (iv(alpha, r_up, nu) - iv(alpha, r_dn, nu)) / (r_up - r_dn)
        d_nu = (iv(alpha, rho, nu + en) - iv(alpha, rho, max(nu - en, 0.0))) / (2.0 * en)  # :274-277
        return d_alpha, d_rho, d_nu


sabr = types.ModuleType(__name__ + ".sabr", "SABR volatility model (B200 drop-in).")
sabr.SABRParameters = SABRParameters
sabr.SABRModel = SABRModel
