"""Batched SABR (Hagan 2002) implied volatilities and smile-calibration objective: thin Python host over
the C ABI (include/heston_b200.h, section SABR; kernels in csrc/sabr_b200.cu).  SURVEY.md 8f rank 4.

``flavour="cpp"`` restates ``SABRModel::implied_volatility`` (src/cpp/models/sabr.cpp:130-192, NaN where the
reference throws), ``flavour="py"`` restates ``SABRCalibrator.sabr_implied_vol``
(calibration/sabr_calibrator.py:159-258), which the calibration objective (:316-324) evaluates.
Parameter rows are ``[alpha, rho, nu]``; ``beta`` is fixed per object as in the reference.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np

from . import _lib
from ._lib import HB_SABR_CPP, HB_SABR_PY, check

_FLAVOURS = {"cpp": HB_SABR_CPP, "py": HB_SABR_PY}
_dp = C.POINTER(C.c_double)


def _np_d(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float64)


class BatchSABR:
    """SABR vols / objective for many parameter sets at once on one CUDA device."""

    def __init__(self, beta: float = 0.5, device: int = 0):
        self._L = _lib.load()
        self.beta, self.device = float(beta), int(device)
        self._smiles = None

    # ---- implied vols ----------------------------------------------------------------------------
    def vols_host(self, params, strikes, forward: float, maturity: float, flavour: str = "py") -> np.ndarray:
        """NumPy in / out: params [P, 3] -> vols [P, n] (copies included)."""
        x = _np_d(np.atleast_2d(params))
        K = _np_d(np.atleast_1d(strikes))
        if x.shape[1] != 3:
            raise ValueError("params rows are (alpha, rho, nu)")
        out = np.empty((x.shape[0], K.size))
        check(self._L.hb_sabr_vols_host(_FLAVOURS[flavour], self.beta, float(forward), float(maturity), K.size,
                                        K.ctypes.data_as(_dp), x.ctypes.data_as(_dp), x.shape[0],
                                        out.ctypes.data_as(_dp)))
        return out

    def vols(self, params, strikes, forward: float, maturity: float, flavour: str = "py"):
        """CUDA tensors: params [P, 3], strikes [n] -> vols [P, n], asynchronously on the current stream."""
        import torch

        if not (isinstance(params, torch.Tensor) and params.is_cuda and params.dtype == torch.float64
                and params.dim() == 2 and params.shape[1] == 3):
            raise ValueError("params must be a CUDA float64 tensor [P, 3]")
        K = torch.as_tensor(strikes, dtype=torch.float64, device=params.device).contiguous()
        soa = params.t().contiguous()
        P, n = soa.shape[1], K.numel()
        out = torch.empty((P, n), dtype=torch.float64, device=params.device)
        with torch.cuda.device(params.device):
            check(self._L.hb_sabr_vols(_FLAVOURS[flavour], self.beta, float(forward), float(maturity), n, K.data_ptr(),
                                       soa.data_ptr(), P, P, out.data_ptr(),
                                       torch.cuda.current_stream(params.device).cuda_stream))
        return out

    # ---- calibration objective over many smiles -----------------------------------------------------
    def set_smiles(self, strikes: Sequence, market_vols: Sequence, forwards, maturities,
                   weights: Optional[Sequence] = None) -> "BatchSABR":
        """Upload the smiles (one per maturity): lists of per-smile arrays.  Weights are normalised per smile
        (sabr_calibrator.py:291-293); default uniform."""
        import torch

        dev = torch.device("cuda", self.device)
        off = np.concatenate([[0], np.cumsum([len(k) for k in strikes])]).astype(np.int32)
        if weights is None:
            weights = [np.ones(len(k)) for k in strikes]
        w = [np.asarray(x, dtype=np.float64) / np.sum(x) for x in weights]
        cat = lambda xs: torch.as_tensor(np.concatenate([_np_d(x) for x in xs]), device=dev)  # noqa: E731
        self._smiles = dict(
            n=len(strikes), off=torch.as_tensor(off, device=dev), max_n=int(np.max(np.diff(off))) if len(strikes) else 0,
            K=cat(strikes), mkt=cat(market_vols), w=cat(w),
            F=torch.as_tensor(_np_d(forwards), device=dev), T=torch.as_tensor(_np_d(maturities), device=dev))
        return self

    def objective(self, params):
        """params: CUDA float64 [n_smiles, P, 3] -> loss [n_smiles, P] = sum_i w_i (sigma_i - market_i)^2."""
        import torch

        s = self._smiles
        if s is None:
            raise RuntimeError("set_smiles has not been called")
        if not (params.is_cuda and params.dtype == torch.float64 and params.dim() == 3 and params.shape[0] == s["n"]
                and params.shape[2] == 3):
            raise ValueError("params must be a CUDA float64 tensor [n_smiles, P, 3]")
        soa = params.permute(0, 2, 1).contiguous()  # [m][3][P]
        P = soa.shape[2]
        out = torch.empty((s["n"], P), dtype=torch.float64, device=params.device)
        with torch.cuda.device(params.device):
            check(self._L.hb_sabr_objective(self.beta, s["n"], s["F"].data_ptr(), s["T"].data_ptr(), s["off"].data_ptr(),
                                            s["max_n"], s["K"].data_ptr(), s["mkt"].data_ptr(), s["w"].data_ptr(),
                                            soa.data_ptr(), P, P, out.data_ptr(),
                                            torch.cuda.current_stream(params.device).cuda_stream))
        return out
