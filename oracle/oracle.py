"""TEST INFRASTRUCTURE ONLY: ctypes loaders for the CPU checkers.

  * ``Oracle()``    -> oracle/liboracle.so       (plain-C restatement, oracle/cpi_oracle.c)
  * ``Reference()`` -> oracle/_ref/libcpi_ref.so (the UNMODIFIED reference compiled in place by oracle/ref_shim.cpp;
                       built only where /root/reference exists, travels prebuilt to the GPU box)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / ``--impl reference`` leg may import this module.
The product package ``cpi_b200`` never does.  Both classes expose the same numpy-level methods.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_P = ctypes.POINTER(ctypes.c_double)
_PI = ctypes.POINTER(ctypes.c_int64)

REC_DOUBLES = {1: 290, 2: 308}
FLAG_IMU_AVG = 1
FLAG_ANALYTIC_JACOBIANS = 2


def build(force: bool = False) -> None:
    """Compile liboracle.so (always possible) and _ref/libcpi_ref.so (only if the reference tree is present)."""
    args = ["make", "-C", _HERE, "-s"]
    if force:
        args.append("-B")
    subprocess.run(args + ["liboracle.so"], check=True)
    subprocess.run(args + ["ref"], check=True)


def _d(a):
    return a.ctypes.data_as(_P)


def _i(a):
    return None if a is None else a.ctypes.data_as(_PI)


class _Base:
    prefix = ""
    path = ""

    def __init__(self):
        if not os.path.exists(self.path):
            raise FileNotFoundError(self.path)
        self.lib = ctypes.CDLL(self.path)
        f = getattr(self.lib, self.prefix + "cpi_preintegrate")
        f.argtypes = [ctypes.c_int, ctypes.c_int64, _PI, ctypes.c_int64, _P, _P, _P, ctypes.c_int, _P, ctypes.c_int]
        f.restype = ctypes.c_int
        self._pre = f
        f = getattr(self.lib, self.prefix + "imu_factor_eval")
        f.argtypes = [ctypes.c_int, ctypes.c_int64, _P, _PI, _PI, _P, _P, _P, _P, _P, ctypes.c_int]
        f.restype = ctypes.c_int
        self._fac = f
        f = getattr(self.lib, self.prefix + "retract")
        f.argtypes = [ctypes.c_int64, _P, _P, _P]
        self._ret = f
        for name, n_in in (("rot_2_quat", 1), ("quat_2_Rot", 1), ("quat_multiply", 2), ("Exp", 1)):
            g = getattr(self.lib, self.prefix + name)
            g.argtypes = [_P] * (n_in + 1)
            g.restype = None

    # ---- batch entry points (same meaning as include/cpi_b200.h, host numpy arrays) ----
    def preintegrate(self, model, samples, lin, sigmas, flags=0, offsets=None, ns=None, nthreads=1):
        samples = np.ascontiguousarray(samples, dtype=np.float64).reshape(-1, 7)
        lin = np.ascontiguousarray(lin, dtype=np.float64).reshape(-1, 13)
        sigmas = np.ascontiguousarray(sigmas, dtype=np.float64)
        n = lin.shape[0]
        if offsets is not None:
            offsets = np.ascontiguousarray(offsets, dtype=np.int64)
            assert offsets.shape[0] == n + 1
            ns = 0
        else:
            extra = 1 if flags & FLAG_IMU_AVG else 0
            if ns is None:
                ns = samples.shape[0] // max(n, 1) - extra
            assert samples.shape[0] >= n * (ns + extra)
        out = np.zeros((n, REC_DOUBLES[model]))
        rc = self._pre(model, n, _i(offsets), int(ns), _d(samples), _d(lin), _d(sigmas), int(flags), _d(out), int(nthreads))
        assert rc == 0
        return out

    def factor_eval(self, model, states, records, lin, idx_i=None, idx_j=None, nthreads=1):
        states = np.ascontiguousarray(states, dtype=np.float64).reshape(-1, 16)
        records = np.ascontiguousarray(records, dtype=np.float64).reshape(-1, REC_DOUBLES[model])
        lin = np.ascontiguousarray(lin, dtype=np.float64).reshape(-1, 13)
        n = records.shape[0]
        if idx_i is not None:
            idx_i = np.ascontiguousarray(idx_i, dtype=np.int64)
            idx_j = np.ascontiguousarray(idx_j, dtype=np.int64)
        e = np.zeros((n, 15)); H1 = np.zeros((n, 225)); H2 = np.zeros((n, 225))
        rc = self._fac(model, n, _d(states), _i(idx_i), _i(idx_j), _d(records), _d(lin), _d(e), _d(H1), _d(H2), int(nthreads))
        assert rc == 0
        return e, H1, H2

    def retract(self, states, xi):
        states = np.ascontiguousarray(states, dtype=np.float64).reshape(-1, 16)
        xi = np.ascontiguousarray(xi, dtype=np.float64).reshape(-1, 15)
        out = np.zeros_like(states)
        self._ret(states.shape[0], _d(states), _d(xi), _d(out))
        return out

    # ---- quat_ops helpers (3x3 as column-major flat [9]) ----
    def _call(self, name, nout, *ins):
        ins = [np.ascontiguousarray(a, dtype=np.float64) for a in ins]
        out = np.zeros(nout)
        getattr(self.lib, self.prefix + name)(*[_d(a) for a in ins], _d(out))
        return out

    def rot_2_quat(self, R): return self._call("rot_2_quat", 4, R)
    def quat_2_Rot(self, q): return self._call("quat_2_Rot", 9, q)
    def quat_multiply(self, q, p): return self._call("quat_multiply", 4, q, p)
    def Exp(self, w): return self._call("Exp", 9, w)


class Oracle(_Base):
    prefix = "oracle_"
    path = os.path.join(_HERE, "liboracle.so")

    def __init__(self):
        super().__init__()
        f = self.lib.oracle_predict_state
        f.argtypes = [ctypes.c_int, ctypes.c_int64, _P, _P, _P, _P]
        self._pred = f

    def predict_state(self, model, states_k, records, lin):
        states_k = np.ascontiguousarray(states_k, dtype=np.float64).reshape(-1, 16)
        records = np.ascontiguousarray(records, dtype=np.float64).reshape(-1, REC_DOUBLES[model])
        lin = np.ascontiguousarray(lin, dtype=np.float64).reshape(-1, 13)
        out = np.zeros_like(states_k)
        self._pred(model, states_k.shape[0], _d(states_k), _d(records), _d(lin), _d(out))
        return out


class Reference(_Base):
    prefix = "ref_"
    path = os.path.join(_HERE, "_ref", "libcpi_ref.so")

    @classmethod
    def available(cls) -> bool:
        return os.path.exists(cls.path)

    def replay_run(self, model, t, w, a, cam_t, lin, sigmas, flags=0, imu_wait=0):
        """ref_replay_run (oracle/ref_shim.cpp): the reference DRIVER over one dataset run with its own deque handling.
        Returns (samples fed [total,7], offsets[nwin+1], records[nwin, rd])."""
        f = self.lib.ref_replay_run
        f.argtypes = [ctypes.c_int, ctypes.c_int64, _P, _P, _P, ctypes.c_int64, _P, ctypes.c_int64, _P, _P, ctypes.c_int, ctypes.c_int64, _P, _PI, _P]
        f.restype = ctypes.c_int64
        t = np.ascontiguousarray(t, dtype=np.float64); w = np.ascontiguousarray(w, dtype=np.float64); a = np.ascontiguousarray(a, dtype=np.float64)
        cam_t = np.ascontiguousarray(cam_t, dtype=np.float64); lin = np.ascontiguousarray(lin, dtype=np.float64)
        sig = np.ascontiguousarray(sigmas, dtype=np.float64)
        rd = 290 if model == 1 else 308
        cap = len(t) + len(cam_t) + 1
        S = np.zeros((cap, 7)); off = np.zeros(len(cam_t) + 1, dtype=np.int64); rec = np.zeros((len(cam_t), rd))
        assert lin.shape[0] >= len(cam_t)
        nw = f(model, len(t), _d(t), _d(w), _d(a), len(cam_t), _d(cam_t), int(imu_wait), _d(lin), _d(sig), flags, cap, _d(S), _i(off), _d(rec))
        if nw < 0:
            raise RuntimeError("ref_replay_run failed")
        return S[:off[nw]].copy(), off[:nw + 1].copy(), rec[:nw].copy()

    def hardware_threads(self) -> int:
        return int(self.lib.ref_hardware_threads())
