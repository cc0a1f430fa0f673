"""ctypes handles onto the parity checkers.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module; nothing under ``pde_b200/`` does.

* :class:`Oracle`  -- ``oracle/liborc.so``: our C restatement (``heston_oracle.c``) of
  /root/reference ``src/cpp/models/heston.cpp:37-167`` and
  ``src/python/quant_trading/calibration/heston_calibrator.py:486-586``.
* :class:`Reference` -- ``oracle/_ref/libheston_ref.so``: the reference's own
  ``heston.cpp`` compiled unmodified behind ``ref_shim.cpp`` (prebuilt here; the GPU
  box has no /root/reference and uses the prebuilt file).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
MODE_REFGRID, MODE_FFT = 0, 1
_dp = C.POINTER(C.c_double)
_u8p = C.POINTER(C.c_uint8)


def _d(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(_dp)


def _b(a):
    a = np.ascontiguousarray(np.asarray(a).astype(np.uint8))
    return a, a.ctypes.data_as(_u8p)


def build(ref: bool = True, pyref: bool = False) -> None:
    """Compile the checkers (gcc/g++ only).  `ref` targets need /root/reference."""
    targets = ["all"]
    if ref and os.path.isdir(os.environ.get("REF", "/root/reference")):
        targets.append("ref")
        if pyref:
            targets.append("pyref")
    subprocess.run(["make", "-C", HERE, *targets], check=True, stdout=subprocess.DEVNULL)


class Oracle:
    """C restatement (port).  FFT mode: CF pinned to the compiled reference, transform / interpolation pinned
    to the independent direct-sum golden tests/golden/fft_direct.npz (see heston_oracle.c)."""

    def __init__(self, path: str | None = None):
        path = path or os.path.join(HERE, "liborc.so")
        if not os.path.exists(path):
            build(ref=False)
        self.lib = L = C.CDLL(path)
        L.orc_price_refgrid.restype = C.c_double
        L.orc_price_refgrid.argtypes = [_dp] + [C.c_double] * 5 + [C.c_int]
        L.orc_objective_from_prices.restype = C.c_double
        L.orc_num_threads.restype = C.c_int

    def num_threads(self) -> int:
        return int(self.lib.orc_num_threads())

    def use_all_cores(self) -> int:
        """OpenMP threads = cores this process may run on (torchrun exports OMP_NUM_THREADS=1)."""
        self.lib.orc_set_num_threads(len(os.sched_getaffinity(0)))
        return self.num_threads()

    def cf(self, p, u, T, S0, r, q) -> complex:
        p_, pp = _d(p)
        out = np.empty(2)
        u = complex(u)
        self.lib.orc_cf(pp, C.c_double(u.real), C.c_double(u.imag), C.c_double(T), C.c_double(S0),
                        C.c_double(r), C.c_double(q), out.ctypes.data_as(_dp))
        return complex(out[0], out[1])

    def cf_grid(self, params, T, ur, ui, S0, r, q) -> np.ndarray:
        """complex128[P][M][n_u] of phi(ur_j + i*ui; T_m) for each parameter set."""
        params, pp = _d(np.atleast_2d(params))
        T, tp = _d(np.atleast_1d(T))
        ur, up = _d(ur)
        out = np.empty((params.shape[0], T.size, ur.size, 2))
        self.lib.orc_cf_grid(C.c_int(params.shape[0]), pp, C.c_int(T.size), tp, C.c_int(ur.size), up,
                             C.c_double(ui), C.c_double(S0), C.c_double(r), C.c_double(q),
                             out.ctypes.data_as(_dp))
        return out[..., 0] + 1j * out[..., 1]

    def cf_ld_grid(self, p, T, ur, ui, S0, r, q) -> np.ndarray:
        """Long-double yardstick (NOT a reference output): complex128[n_u]."""
        p_, pp = _d(p)
        ur, up = _d(ur)
        out = np.empty((ur.size, 2))
        self.lib.orc_cf_ld_grid(pp, C.c_int(ur.size), up, C.c_double(ui), C.c_double(T), C.c_double(S0),
                                C.c_double(r), C.c_double(q), out.ctypes.data_as(_dp))
        return out[:, 0] + 1j * out[:, 1]

    def price_refgrid(self, p, K, T, S0, r, q, is_call=True) -> float:
        p_, pp = _d(p)
        return float(self.lib.orc_price_refgrid(pp, K, T, S0, r, q, int(bool(is_call))))

    def fft_slice(self, p, T, S0, r, q, N=4096, eta=0.25, alpha=0.75) -> np.ndarray:
        p_, pp = _d(p)
        grid = np.empty(N)
        work = np.empty(2 * N)
        self.lib.orc_fft_slice(pp, C.c_double(T), C.c_double(S0), C.c_double(r), C.c_double(q), C.c_int(N),
                               C.c_double(eta), C.c_double(alpha), grid.ctypes.data_as(_dp),
                               work.ctypes.data_as(C.c_void_p))
        return grid

    def _surf(self, K, T, is_call):
        K, kp = _d(K)
        T = np.broadcast_to(np.asarray(T, dtype=np.float64), K.shape)
        T, tp = _d(T)
        ic = np.broadcast_to(np.asarray(is_call), K.shape)
        ic, ip = _b(ic)
        return (K, kp), (T, tp), (ic, ip)

    def price_batch(self, mode, params, K, T, is_call, S0, r, q, N=4096, eta=0.25, alpha=0.75) -> np.ndarray:
        params, pp = _d(np.atleast_2d(params))
        (K, kp), (T, tp), (ic, ip) = self._surf(K, T, is_call)
        out = np.empty((params.shape[0], K.size))
        self.lib.orc_price_batch(C.c_int(mode), C.c_int(params.shape[0]), pp, C.c_int(K.size), kp, tp, ip,
                                 C.c_double(S0), C.c_double(r), C.c_double(q), C.c_int(N), C.c_double(eta),
                                 C.c_double(alpha), out.ctypes.data_as(_dp))
        return out

    def objective_batch(self, mode, params, K, T, is_call, market, S0, r, q, N=4096, eta=0.25,
                        alpha=0.75) -> np.ndarray:
        params, pp = _d(np.atleast_2d(params))
        (K, kp), (T, tp), (ic, ip) = self._surf(K, T, is_call)
        mk, mp = _d(market)
        out = np.empty(params.shape[0])
        self.lib.orc_objective_batch(C.c_int(mode), C.c_int(params.shape[0]), pp, C.c_int(K.size), kp, tp, ip, mp,
                                     C.c_double(S0), C.c_double(r), C.c_double(q), C.c_int(N), C.c_double(eta),
                                     C.c_double(alpha), out.ctypes.data_as(_dp))
        return out

    def residuals_from_prices(self, prices, market) -> np.ndarray:
        pr, pp = _d(prices)
        mk, mp = _d(market)
        out = np.empty(pr.size)
        self.lib.orc_residuals_from_prices(C.c_int(pr.size), pp, mp, out.ctypes.data_as(_dp))
        return out

    def objective_from_prices(self, prices, market) -> float:
        pr, pp = _d(prices)
        mk, mp = _d(market)
        return float(self.lib.orc_objective_from_prices(C.c_int(pr.size), pp, mp))

    def fd_steps(self, x, lb, ub) -> np.ndarray:
        x, xp = _d(x)
        lb, lp = _d(lb)
        ub, up = _d(ub)
        h = np.empty(5)
        self.lib.orc_fd_steps(xp, lp, up, h.ctypes.data_as(_dp))
        return h

    def jacobian(self, mode, p, lb, ub, K, T, is_call, market, S0, r, q, N=4096, eta=0.25, alpha=0.75):
        p, pp = _d(p)
        lb, lp = _d(lb)
        ub, up = _d(ub)
        (K, kp), (T, tp), (ic, ip) = self._surf(K, T, is_call)
        mk, mp = _d(market)
        r0 = np.empty(K.size)
        J = np.empty((K.size, 5))
        self.lib.orc_jacobian(C.c_int(mode), pp, lp, up, C.c_int(K.size), kp, tp, ip, mp, C.c_double(S0),
                              C.c_double(r), C.c_double(q), C.c_int(N), C.c_double(eta), C.c_double(alpha),
                              r0.ctypes.data_as(_dp), J.ctypes.data_as(_dp))
        return r0, J

    def normal_eq_batch(self, mode, params, lb, ub, K, T, is_call, market, S0, r, q, N=4096, eta=0.25,
                        alpha=0.75) -> np.ndarray:
        params, pp = _d(np.atleast_2d(params))
        lb, lp = _d(lb)
        ub, up = _d(ub)
        (K, kp), (T, tp), (ic, ip) = self._surf(K, T, is_call)
        mk, mp = _d(market)
        out = np.empty((params.shape[0], 22))
        self.lib.orc_normal_eq_batch(C.c_int(mode), C.c_int(params.shape[0]), pp, lp, up, C.c_int(K.size), kp, tp,
                                     ip, mp, C.c_double(S0), C.c_double(r), C.c_double(q), C.c_int(N),
                                     C.c_double(eta), C.c_double(alpha), out.ctypes.data_as(_dp))
        return out


def _sabr_args(K, params):
    K_, kp = _d(np.atleast_1d(K))
    x_, xp = _d(np.atleast_2d(params))
    if x_.shape[1] != 3:
        raise ValueError("SABR parameter rows are (alpha, rho, nu)")
    return K_, kp, x_, xp


class SabrOracle:
    """C restatement (oracle/sabr_oracle.c) of the reference's two SABR formulas: flavour "cpp" =
    SABRModel::implied_volatility (sabr.cpp:130-192), "py" = SABRCalibrator.sabr_implied_vol
    (sabr_calibrator.py:159-258) and the calibration objective (:316-324)."""

    def __init__(self, path: str | None = None):
        path = path or os.path.join(HERE, "liborc.so")
        if not os.path.exists(path):
            build(ref=False)
        self.lib = C.CDLL(path)

    def vols(self, flavour: str, beta, F, T, K, params) -> np.ndarray:
        K_, kp, x_, xp = _sabr_args(K, params)
        out = np.empty((x_.shape[0], K_.size))
        self.lib.orc_sabr_vols(C.c_int({"cpp": 0, "py": 1}[flavour]), C.c_double(beta), C.c_double(F), C.c_double(T),
                               C.c_int(K_.size), kp, C.c_int(x_.shape[0]), xp, out.ctypes.data_as(_dp))
        return out

    def objective(self, beta, F, T, K, market, weights, params) -> np.ndarray:
        K_, kp, x_, xp = _sabr_args(K, params)
        m_, mp = _d(market)
        w_, wp = _d(weights)
        out = np.empty(x_.shape[0])
        self.lib.orc_sabr_objective(C.c_double(beta), C.c_double(F), C.c_double(T), C.c_int(K_.size), kp, mp, wp,
                                    C.c_int(x_.shape[0]), xp, out.ctypes.data_as(_dp))
        return out


class Reference:
    """The reference's own heston.cpp (unmodified) behind oracle/ref_shim.cpp."""

    def __init__(self, path: str | None = None):
        path = path or os.path.join(HERE, "_ref", "libheston_ref.so")
        if not os.path.exists(path):
            build(ref=True)
        if not os.path.exists(path):
            raise FileNotFoundError(
                f"{path} missing and /root/reference not present to build it (run `make -C oracle ref` "
                "in the build container; the file travels with the gpurun snapshot)")
        self.lib = L = C.CDLL(path)
        L.ref_last_error.restype = C.c_char_p
        L.ref_num_threads.restype = C.c_int

    def last_error(self) -> str:
        return self.lib.ref_last_error().decode()

    def num_threads(self) -> int:
        return int(self.lib.ref_num_threads())

    def use_all_cores(self) -> int:
        """OpenMP threads = cores this process may run on (torchrun exports OMP_NUM_THREADS=1)."""
        self.lib.ref_set_num_threads(len(os.sched_getaffinity(0)))
        return self.num_threads()

    def _chk(self, rc):
        if rc:
            raise ValueError(self.last_error())

    def validate(self, p) -> None:
        p_, pp = _d(p)
        self._chk(self.lib.ref_validate(pp))

    def cf(self, p, u, T, S0, r, q) -> complex:
        p_, pp = _d(p)
        u = complex(u)
        out = np.empty(2)
        self._chk(self.lib.ref_cf(pp, C.c_double(u.real), C.c_double(u.imag), C.c_double(T), C.c_double(S0),
                                  C.c_double(r), C.c_double(q), out.ctypes.data_as(_dp)))
        return complex(out[0], out[1])

    def cf_grid(self, p, ur, ui, T, S0, r, q) -> np.ndarray:
        p_, pp = _d(p)
        ur, up = _d(ur)
        out = np.empty((ur.size, 2))
        self._chk(self.lib.ref_cf_grid(pp, C.c_int(ur.size), up, C.c_double(ui), C.c_double(T), C.c_double(S0),
                                       C.c_double(r), C.c_double(q), out.ctypes.data_as(_dp)))
        return out[:, 0] + 1j * out[:, 1]

    def price_option(self, p, K, T, S0, r, q, is_call=True) -> float:
        p_, pp = _d(p)
        out = C.c_double()
        self._chk(self.lib.ref_price_option(pp, C.c_double(K), C.c_double(T), C.c_double(S0), C.c_double(r),
                                            C.c_double(q), int(bool(is_call)), C.byref(out)))
        return out.value

    def price_options(self, p, K, T, S0, r, q, is_call=True) -> np.ndarray:
        p_, pp = _d(p)
        K, kp = _d(K)
        T, tp = _d(np.atleast_1d(T))
        out = np.empty(K.size)
        self._chk(self.lib.ref_price_options(pp, C.c_int(K.size), kp, C.c_int(T.size), tp, C.c_double(S0),
                                             C.c_double(r), C.c_double(q), int(bool(is_call)),
                                             out.ctypes.data_as(_dp)))
        return out

    def price_surface_batch(self, params, K, T, S0, r, q) -> np.ndarray:
        params, pp = _d(np.atleast_2d(params))
        K, kp = _d(K)
        T, tp = _d(np.broadcast_to(np.asarray(T, dtype=np.float64), K.shape))
        out = np.empty((params.shape[0], K.size))
        self.lib.ref_price_surface_batch(C.c_int(params.shape[0]), pp, C.c_int(K.size), kp, tp, C.c_double(S0),
                                         C.c_double(r), C.c_double(q), out.ctypes.data_as(_dp))
        return out

    def implied_vol(self, p, K, T, S0, r, q, is_call=True) -> float:
        p_, pp = _d(p)
        out = C.c_double()
        self._chk(self.lib.ref_implied_vol(pp, C.c_double(K), C.c_double(T), C.c_double(S0), C.c_double(r),
                                           C.c_double(q), int(bool(is_call)), C.byref(out)))
        return out.value

    def sabr_vols(self, beta, F, T, K, params) -> np.ndarray:
        """The reference's SABRModel::implied_volatility (sabr.cpp compiled unmodified); NaN where it throws."""
        K_, kp, x_, xp = _sabr_args(K, params)
        out = np.empty((x_.shape[0], K_.size))
        self.lib.ref_sabr_vols(C.c_double(beta), C.c_double(F), C.c_double(T), C.c_int(K_.size), kp,
                               C.c_int(x_.shape[0]), xp, out.ctypes.data_as(_dp))
        return out

    def greeks(self, p, K, T, S0, r, q, is_call=True) -> np.ndarray:
        p_, pp = _d(p)
        out = np.empty(6)
        self._chk(self.lib.ref_greeks(pp, C.c_double(K), C.c_double(T), C.c_double(S0), C.c_double(r),
                                      C.c_double(q), int(bool(is_call)), out.ctypes.data_as(_dp)))
        return out
