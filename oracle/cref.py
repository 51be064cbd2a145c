"""Build + ctypes binding of ``oracle/ball_sim_ref.c`` (the plain-C restatement).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Built artefacts go to
``oracle/_build/`` (git-ignored ``*.so``).  Two flavours:
  * ``portable``: ``-O2`` baseline x86-64, built by ``__graft_entry__.build()`` here
    and shipped to the GPU box with the snapshot — the parity checker.
  * ``native``: ``-O3 -march=native``, rebuilt on whatever box runs ``bench.py`` —
    the timed CPU baseline.  Both use ``-ffp-contract=off`` and no fast-math, so
    they are bit-identical to each other and to ``ball_sim_np.py``.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_SRC = _HERE / "ball_sim_ref.c"
_INC = _HERE / "ball_sim_step.inc"
_OUT = _HERE / "_build"

_FLAGS = {
    "portable": ["-O2"],
    "native": ["-O3", "-march=native"],
}


_POLICY_H = _HERE.parent / "include" / "doper_fp_policy.h"

# bits of include/doper_fp_policy.h (which mul+add pairs are single-rounding FMAs)
FP_POS, FP_VEL, FP_LERP, FP_DOT, FP_DOT_REV, FP_FORCE = 1, 2, 4, 8, 16, 32
FP_DEFAULT = FP_POS


def build(flavour: str = "portable", force: bool = False, policy: int | None = None) -> Path:
    """``policy``: contraction mask (None = the header's default, shared with the product kernels)."""
    _OUT.mkdir(exist_ok=True)
    so = _OUT / (f"liboracle_{flavour}.so" if policy is None else f"liboracle_{flavour}_p{int(policy)}.so")
    newest = max(_SRC.stat().st_mtime, _INC.stat().st_mtime, _POLICY_H.stat().st_mtime)
    if so.exists() and not force and so.stat().st_mtime >= newest:
        return so
    cmd = ["gcc", "-std=c11", "-shared", "-fPIC", "-fopenmp", "-ffp-contract=off", "-fno-fast-math",
           *_FLAGS[flavour], *([] if policy is None else [f"-DDOPER_FP_POLICY={int(policy)}"]),
           str(_SRC), "-o", str(so), "-lm"]
    subprocess.run(cmd, check=True)
    return so


_f = ctypes.POINTER(ctypes.c_float)
_i = ctypes.POINTER(ctypes.c_int32)
_u8 = ctypes.POINTER(ctypes.c_uint8)


def _fp(a):
    return None if a is None else a.ctypes.data_as(_f)


class CRef:
    """Thin numpy wrapper. All arrays are float32 C-contiguous."""

    def __init__(self, flavour: str = "portable", policy: int | None = None):
        self.lib = ctypes.CDLL(str(build(flavour, policy=policy)))
        L = self.lib
        L.oracle_fp_policy.restype = ctypes.c_int
        self.policy = int(L.oracle_fp_policy())
        L.oracle_closest_segment.argtypes = [ctypes.c_int64, ctypes.c_int, _f, _f, _i, _f]
        L.oracle_points_inside.argtypes = [ctypes.c_int64, ctypes.c_int, _f, _f, _u8]
        L.oracle_rollout_fwd.argtypes = [ctypes.c_int64, ctypes.c_int, ctypes.c_int64, ctypes.c_int,
                                         ctypes.c_float, _f, _f, _f, _f, _f, _f, _f, _f, _f, _f, _i]
        L.oracle_value_and_grad.argtypes = [ctypes.c_int64, ctypes.c_int, ctypes.c_int64, ctypes.c_int,
                                            ctypes.c_float, _f, _f, _f, _f, _f, _f, _f, _f, _f, _f, _f, _i]
        _d = ctypes.POINTER(ctypes.c_double)
        L.oracle_grad_truth64.argtypes = [ctypes.c_int64, ctypes.c_int, ctypes.c_int64, ctypes.c_int,
                                          ctypes.c_float, _f, _f, _f, _f, _f, _f, _d, _d]
        for fn in (L.oracle_closest_segment, L.oracle_points_inside, L.oracle_rollout_fwd,
                   L.oracle_value_and_grad, L.oracle_grad_truth64):
            fn.restype = None

    @staticmethod
    def set_threads(n: int):
        os.environ["OMP_NUM_THREADS"] = str(n)

    @staticmethod
    def _c(a):
        return np.ascontiguousarray(a, dtype=np.float32)

    @staticmethod
    def consts_array(c) -> np.ndarray:
        return np.array([c.radius, c.mass, c.friction, c.elasticity, c.k_s], dtype=np.float32)

    def closest_segment(self, points, segments):
        p = self._c(points).reshape(-1, 2)
        s = self._c(segments).reshape(-1, 4)
        idx = np.empty(p.shape[0], np.int32)
        dist = np.empty(p.shape[0], np.float32)
        self.lib.oracle_closest_segment(p.shape[0], s.shape[0], _fp(p), _fp(s),
                                        idx.ctypes.data_as(_i), _fp(dist))
        return idx, dist

    def points_inside(self, points, segments):
        p = self._c(points).reshape(-1, 2)
        s = self._c(segments).reshape(-1, 4)
        flag = np.empty(p.shape[0], np.uint8)
        self.lib.oracle_points_inside(p.shape[0], s.shape[0] // 3, _fp(p), _fp(s),
                                      flag.ctypes.data_as(_u8))
        return flag.astype(bool)

    def _scene(self, segments):
        s = self._c(segments)
        if s.ndim == 4:  # per-env [B,S,2,2]
            return s.reshape(s.shape[0], -1, 4), s.shape[1], s.shape[1] * 4
        s = s.reshape(-1, 4)
        return s, s.shape[0], 0

    def rollout_fwd(self, n_steps, dt, consts, attractor, segments, x0, v0, want_traj=False):
        s, S, stride = self._scene(segments)
        x0 = self._c(x0).reshape(-1, 2)
        v0 = self._c(v0).reshape(-1, 2)
        B = x0.shape[0]
        xT, vT = np.empty_like(x0), np.empty_like(v0)
        hl = np.empty((n_steps, B), np.int32)
        tx = tv = ta = None
        if want_traj:
            tx, tv, ta = (np.empty((n_steps, B, 2), np.float32) for _ in range(3))
        k = self.consts_array(consts)
        A = self._c(attractor).reshape(2)
        self.lib.oracle_rollout_fwd(B, S, stride, n_steps, np.float32(dt), _fp(k), _fp(A), _fp(s),
                                    _fp(x0), _fp(v0), _fp(xT), _fp(vT), _fp(tx), _fp(tv), _fp(ta),
                                    hl.ctypes.data_as(_i))
        return {"x_T": xT, "v_T": vT, "traj_x": tx, "traj_v": tv, "traj_a": ta,
                "idx": hl & 0x7FFFFFFF, "hit": hl < 0, "hitlog": hl}

    def value_and_grad(self, n_steps, dt, consts, attractor, target, segments, x0, v0):
        s, S, stride = self._scene(segments)
        x0 = self._c(x0).reshape(-1, 2)
        v0 = self._c(v0).reshape(-1, 2)
        B = x0.shape[0]
        xT, vT, gv, gx = (np.empty_like(x0) for _ in range(4))
        loss = np.empty(B, np.float32)
        hl = np.empty((n_steps, B), np.int32)
        k = self.consts_array(consts)
        A = self._c(attractor).reshape(2)
        tg = self._c(target).reshape(2)
        self.lib.oracle_value_and_grad(B, S, stride, n_steps, np.float32(dt), _fp(k), _fp(A), _fp(tg),
                                       _fp(s), _fp(x0), _fp(v0), _fp(loss), _fp(xT), _fp(vT), _fp(gv),
                                       _fp(gx), hl.ctypes.data_as(_i))
        return {"loss_per_ball": loss, "x_T": xT, "v_T": vT, "grad_v0": gv, "grad_x0": gx,
                "idx": hl & 0x7FFFFFFF, "hit": hl < 0, "hitlog": hl}

    def grad_truth64(self, n_steps, dt, consts, attractor, target, segments, x0, v0):
        """float64 adjoint on the fp32 forward history -> (dL/dv0, dL/dx0) as float64 [B,2]."""
        s, S, stride = self._scene(segments)
        x0 = self._c(x0).reshape(-1, 2)
        v0 = self._c(v0).reshape(-1, 2)
        B = x0.shape[0]
        gv, gx = np.empty((B, 2), np.float64), np.empty((B, 2), np.float64)
        k = self.consts_array(consts)
        A = self._c(attractor).reshape(2)
        tg = self._c(target).reshape(2)
        dp = ctypes.POINTER(ctypes.c_double)
        self.lib.oracle_grad_truth64(B, S, stride, n_steps, np.float32(dt), _fp(k), _fp(A), _fp(tg), _fp(s),
                                     _fp(x0), _fp(v0), gv.ctypes.data_as(dp), gx.ctypes.data_as(dp))
        return gv, gx
