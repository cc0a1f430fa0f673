"""Batched Heston pricer: thin Python host over the C ABI (include/heston_b200.h).

PyTorch is used for device buffers and streams only; all arithmetic happens in
libheston_b200.so.  One :class:`BatchPricer` = one ``hb_plan`` (grid constants + one option
surface).  Batched counterparts of the reference's per-candidate calls:

=====================  ==============================================================
``price``              HestonCalibrator._price_options   (heston_calibrator.py:538-586)
``objective``          HestonCalibrator._compute_objective (heston_calibrator.py:486-513)
``jacobian``           _compute_residuals + SciPy 2-point Jacobian (heston_calibrator.py:515-536)
``normal_equations``   loss, ||r||^2, J^T r, triu(J^T J) per parameter set
``implied_vol``        HestonModel::implied_volatility   (heston.cpp:311-349)
``greeks``             HestonModel::price_option_with_greeks (heston.cpp:168-217)
``characteristic_function``  HestonModel::characteristic_function (heston.cpp:74-92)
=====================  ==============================================================

Parameter sets are rows ``[kappa, theta, sigma, rho, v0]``.  Methods taking a CUDA
``torch.Tensor`` run asynchronously on the current stream and return CUDA tensors; the
``*_host`` methods take/return NumPy arrays and include the host<->device copies.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np

from . import _lib
from ._lib import HB_MODE_FFT, HB_MODE_REFGRID, HB_NEQ_WIDTH, check

_MODES = {"refgrid": HB_MODE_REFGRID, "fft": HB_MODE_FFT}
_dp = C.POINTER(C.c_double)
_u8p = C.POINTER(C.c_uint8)


def _np_d(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float64)


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(_dp)


class BatchPricer:
    """Owns one ``hb_plan``.

    Args:
        mode: ``"fft"`` (Carr-Madan ``n_grid``-point FFT, linear log-strike interpolation) or
            ``"refgrid"`` (the reference's 1023-point quadrature, heston.cpp:94-151 -- what
            ``quant_cpp.heston.HestonModel.price_option`` returns).
        n_grid, eta, alpha: FFT grid (reference docs: 4096, 0.25; alpha = 0.75, heston.hpp:261).
        device: CUDA ordinal.
    """

    def __init__(self, mode: str = "fft", n_grid: int = 4096, eta: float = 0.25, alpha: float = 0.75,
                 device: int = 0):
        if mode not in _MODES:
            raise ValueError(f"mode must be 'fft' or 'refgrid', got {mode!r}")
        self._L = _lib.load()
        self.mode, self.n_grid, self.eta, self.alpha, self.device = mode, int(n_grid), float(eta), float(alpha), int(device)
        h = C.c_void_p()
        check(self._L.hb_plan_create(_MODES[mode], self.n_grid, self.eta, self.alpha, self.device, C.byref(h)))
        self._h = h
        self.n_options = 0
        self.n_maturities = 0

    def close(self) -> None:
        if getattr(self, "_h", None):
            self._L.hb_plan_destroy(self._h)
            self._h = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    # ---- surface ------------------------------------------------------------------------------
    def set_surface(self, strikes, maturities, is_calls=True, market: Optional[Sequence[float]] = None, *,
                    S0: float, r: float, q: float) -> "BatchPricer":
        """Flat option list as the calibrator holds it (heston_calibrator.py:293-303).  `maturities`
        and `is_calls` broadcast against `strikes`."""
        K = _np_d(np.atleast_1d(strikes))
        T = _np_d(np.broadcast_to(np.asarray(maturities, dtype=np.float64), K.shape))
        ic = np.ascontiguousarray(np.broadcast_to(np.asarray(is_calls), K.shape).astype(np.uint8))
        mk = None
        if market is not None:
            mk = _np_d(np.atleast_1d(market))
            if mk.shape != K.shape:
                raise ValueError("market must have the same length as strikes")
        check(self._L.hb_surface_set(self._h, K.size, _ptr(K), _ptr(T), ic.ctypes.data_as(_u8p),
                                     _ptr(mk) if mk is not None else None, float(S0), float(r), float(q)))
        self.n_options = K.size
        self.n_maturities = int(self._L.hb_plan_n_maturities(self._h))
        return self

    def set_truncation(self, abs_price_error: float) -> "BatchPricer":
        """Admissible absolute price error of the significance cut (``hb_plan_set_truncation``): grid points
        whose |phi| is so small that all of them together cannot move a price by more than this are treated as
        exact zeros.  Default 2**-80; 0.0 = exact mode (only true exp underflow is skipped)."""
        check(self._L.hb_plan_set_truncation(self._h, float(abs_price_error)))
        return self

    @property
    def log_cut(self) -> float:
        """log|phi| below which a grid point counts as 0 for the surface in force (-746 = exp underflow)."""
        return float(self._L.hb_plan_log_cut(self._h))

    def profile(self, enable: bool = True) -> "BatchPricer":
        """Record CUDA-event pairs around the kernels of every pricing launch (``hb_plan_profile``)."""
        check(self._L.hb_plan_profile(self._h, 1 if enable else 0))
        return self

    def profile_read(self) -> dict:
        """Summed kernel durations since the last read (``hb_plan_profile_read``): prefix scan, direct-sum job kernel,
        transform / refgrid job kernel; how the last launch was routed.  Waits for the launches."""
        ms = (C.c_double * 3)()
        n = (C.c_longlong * 3)()
        check(self._L.hb_plan_profile_read(self._h, ms, n))
        return {"scan_ms": ms[0], "direct_ms": ms[1], "transform_ms": ms[2], "kernel_invocations": int(n[0]),
                "sets_direct": int(n[1]), "sets_transform": int(n[2])}

    def set_bounds(self, lb, ub) -> "BatchPricer":
        lb, ub = _np_d(lb), _np_d(ub)
        if lb.size != 5 or ub.size != 5:
            raise ValueError("bounds need 5 entries: kappa, theta, sigma, rho, v0")
        check(self._L.hb_set_bounds(self._h, _ptr(lb), _ptr(ub)))
        return self

    # ---- device-tensor API ------------------------------------------------------------------------
    def _soa(self, params):
        import torch

        if not (isinstance(params, torch.Tensor) and params.is_cuda):
            raise TypeError("expected a CUDA torch.Tensor of shape [P, 5] (use the *_host methods for NumPy)")
        if params.dtype != torch.float64 or params.dim() != 2 or params.shape[1] != 5:
            raise ValueError("params must be float64 with shape [P, 5]")
        if params.device.index != self.device:
            raise ValueError(f"params live on cuda:{params.device.index}, plan on cuda:{self.device}")
        return params.t().contiguous(), torch.cuda.current_stream(params.device).cuda_stream

    def price(self, params):
        """[P, 5] -> prices [P, n_options] (NaN rows for invalid parameter sets)."""
        import torch

        soa, st = self._soa(params)
        P = soa.shape[1]
        out = torch.empty((P, self.n_options), dtype=torch.float64, device=params.device)
        check(self._L.hb_price(self._h, soa.data_ptr(), P, P, out.data_ptr(), st))
        return out

    def objective(self, params):
        """[P, 5] -> loss [P]: sum of squared relative errors, 1e10 if any price is NaN or <= 0."""
        import torch

        soa, st = self._soa(params)
        P = soa.shape[1]
        out = torch.empty((P,), dtype=torch.float64, device=params.device)
        check(self._L.hb_objective(self._h, soa.data_ptr(), P, P, out.data_ptr(), st))
        return out

    def implied_vol(self, params):
        """[P, 5] -> Black-Scholes implied vols of the model prices [P, n_options]
        (batched HestonModel::implied_volatility, heston.cpp:311-349)."""
        import torch

        soa, st = self._soa(params)
        P = soa.shape[1]
        out = torch.empty((P, self.n_options), dtype=torch.float64, device=params.device)
        check(self._L.hb_implied_vol(self._h, soa.data_ptr(), P, P, out.data_ptr(), st))
        return out

    def greeks(self, params):
        """[P, 5] -> finite-difference Greeks [P, n_options, 5] = delta, gamma, vega, theta, rho
        (batched HestonModel::price_option_with_greeks, heston.cpp:168-217: same bumps, same formulas)."""
        import torch

        soa, st = self._soa(params)
        P = soa.shape[1]
        out = torch.empty((P, self.n_options, 5), dtype=torch.float64, device=params.device)
        check(self._L.hb_greeks(self._h, soa.data_ptr(), P, P, out.data_ptr(), st))
        return out

    def normal_equations(self, params):
        """[P, 5] -> [P, 22] = loss, ||r||^2, J^T r (5), triu(J^T J) (15)."""
        import torch

        soa, st = self._soa(params)
        P = soa.shape[1]
        out = torch.empty((P, HB_NEQ_WIDTH), dtype=torch.float64, device=params.device)
        check(self._L.hb_normal_eq(self._h, soa.data_ptr(), P, P, out.data_ptr(), st))
        return out

    def jacobian(self, params):
        """[P, 5] -> (residuals [P, n], jacobian [P, n, 5]) with SciPy's 2-point step rule."""
        import torch

        soa, st = self._soa(params)
        P = soa.shape[1]
        res = torch.empty((P, self.n_options), dtype=torch.float64, device=params.device)
        jac = torch.empty((P, self.n_options, 5), dtype=torch.float64, device=params.device)
        check(self._L.hb_jacobian(self._h, soa.data_ptr(), P, P, res.data_ptr(), jac.data_ptr(), st))
        return res, jac

    # ---- host (NumPy) API: copies included ----------------------------------------------------------
    def price_host(self, params) -> np.ndarray:
        x = _np_d(np.atleast_2d(params))
        out = np.empty((x.shape[0], self.n_options))
        check(self._L.hb_price_host(self._h, _ptr(x), x.shape[0], _ptr(out)))
        return out

    def implied_vol_host(self, params) -> np.ndarray:
        x = _np_d(np.atleast_2d(params))
        out = np.empty((x.shape[0], self.n_options))
        check(self._L.hb_implied_vol_host(self._h, _ptr(x), x.shape[0], _ptr(out)))
        return out

    def greeks_host(self, params) -> np.ndarray:
        x = _np_d(np.atleast_2d(params))
        out = np.empty((x.shape[0], self.n_options, 5))
        check(self._L.hb_greeks_host(self._h, _ptr(x), x.shape[0], _ptr(out)))
        return out

    def objective_host(self, params) -> np.ndarray:
        x = _np_d(np.atleast_2d(params))
        out = np.empty(x.shape[0])
        check(self._L.hb_objective_host(self._h, _ptr(x), x.shape[0], _ptr(out)))
        return out

    def normal_equations_host(self, params) -> np.ndarray:
        x = _np_d(np.atleast_2d(params))
        out = np.empty((x.shape[0], HB_NEQ_WIDTH))
        check(self._L.hb_normal_eq_host(self._h, _ptr(x), x.shape[0], _ptr(out)))
        return out

    def jacobian_host(self, params):
        x = _np_d(np.atleast_2d(params))
        res = np.empty((x.shape[0], self.n_options))
        jac = np.empty((x.shape[0], self.n_options, 5))
        check(self._L.hb_jacobian_host(self._h, _ptr(x), x.shape[0], _ptr(res), _ptr(jac)))
        return res, jac


def characteristic_function(params, T, u, *, S0: float, r: float, q: float):
    """phi(u_j; T_m) for every parameter set: CUDA tensors in ([P,5], [M], complex [n_u]) ->
    complex128 [P, M, n_u].  Batched HestonModel::characteristic_function (heston.cpp:74-92)."""
    import torch

    L = _lib.load()
    if not params.is_cuda:
        raise TypeError("expected CUDA tensors")
    dev = params.device
    soa = params.to(torch.float64).t().contiguous()
    T = torch.as_tensor(T, dtype=torch.float64, device=dev).contiguous()
    u = torch.as_tensor(u, dtype=torch.complex128, device=dev)
    ur, ui = u.real.contiguous(), u.imag.contiguous()
    P, M, n_u = soa.shape[1], T.numel(), u.numel()
    out = torch.empty((P, M, n_u), dtype=torch.complex128, device=dev)
    with torch.cuda.device(dev):
        check(L.hb_cf(soa.data_ptr(), P, P, T.data_ptr(), M, ur.data_ptr(), ui.data_ptr(), n_u, float(S0), float(r),
                      float(q), out.data_ptr(), torch.cuda.current_stream(dev).cuda_stream))
    return out


def fft_batch(x):
    """In-place forward FFT of a CUDA complex128 tensor [n_slices, n] (n = 512 or 4096) through the
    shared-memory Stockham kernel.  Returns `x`."""
    import torch

    L = _lib.load()
    if not (x.is_cuda and x.dtype == torch.complex128 and x.dim() == 2 and x.is_contiguous()):
        raise ValueError("expected a contiguous CUDA complex128 tensor [n_slices, n]")
    with torch.cuda.device(x.device):
        check(L.hb_fft_batch(x.data_ptr(), x.shape[1], x.shape[0], torch.cuda.current_stream(x.device).cuda_stream))
    return x


def measure_fp64_peak(device: int = 0, seconds: float = 0.3) -> float:
    """Sustained DFMA TFLOP/s of `device` (roofline denominator of this FP64-bound path)."""
    L = _lib.load()
    out = C.c_double()
    check(L.hb_measure_fp64_peak(int(device), float(seconds), C.byref(out)))
    return out.value


def launch_count() -> int:
    return int(_lib.load().hb_launch_count())
