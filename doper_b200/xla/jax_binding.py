"""jax.ffi + jax.custom_vjp binding of the C ABI  —  UNTESTED (no jax / jaxlib in this image).

This is the boundary BASELINE.json's north_star names.  It cannot be imported here
(``import jax`` raises) and has never run; it is kept as reviewed source so that an image with
``jaxlib`` can build ``xla_ffi_shim.cc`` (against ``jax.ffi.include_dir()``) and get the same
kernels behind ``jax.custom_vjp``.  All tested paths use ``doper_b200.sim`` (torch / ctypes).

    from doper_b200.xla.jax_binding import rollout      # needs jax
    x_T, v_T = rollout(scene_ws, x0, v0, S=201, n_steps=300, dt=1e-3, constants=..., attractor=...)
    jax.grad(lambda v: jnp.abs(rollout(ws, x0, v, ...)[0] - target).sum())(v0)
"""
from __future__ import annotations

import ctypes
import subprocess
from functools import partial
from pathlib import Path

import numpy as np

try:  # pragma: no cover - jax is absent in the build image
    import jax
    import jax.numpy as jnp
except ImportError as e:  # pragma: no cover
    raise ImportError("doper_b200.xla.jax_binding needs jax/jaxlib, which this image does not provide; "
                      "use doper_b200.sim (torch/ctypes host binding over the same C ABI)") from e

from .._abi import LIB_PATH, RolloutDesc, lib  # pragma: no cover

_HERE = Path(__file__).resolve().parent
_SHIM = _HERE / "libdoper_xla_ffi.so"


def _build_shim() -> Path:  # pragma: no cover
    if not _SHIM.exists():
        subprocess.run(["nvcc", "-std=c++17", "-O2", "-shared", "-Xcompiler", "-fPIC",
                        f"-I{jax.ffi.include_dir()}", str(_HERE / "xla_ffi_shim.cc"), "-o", str(_SHIM),
                        f"-L{LIB_PATH.parent}", "-ldoper_b200", f"-Xlinker=-rpath={LIB_PATH.parent}"], check=True)
    return _SHIM


def _register():  # pragma: no cover
    shim = ctypes.CDLL(str(_build_shim()))
    for name in ("doper_scene_prepare_ffi", "doper_rollout_fwd_ffi", "doper_rollout_bwd_ffi"):
        jax.ffi.register_ffi_target(name, jax.ffi.pycapsule(getattr(shim, name)), platform="CUDA")


_register()  # pragma: no cover


def _attrs(S, per_env, n_steps, ckpt_every, dt, constants, attractor):  # pragma: no cover
    f = np.float32
    return dict(S=np.int32(S), per_env=np.int32(per_env), n_steps=np.int32(n_steps), ckpt_every=np.int32(ckpt_every),
                dt=f(dt), radius=f(constants["radius"]), mass=f(constants["mass"]),
                friction=f(constants["rolling_friction_coefficient"]), elasticity=f(constants["walls_elasticity"]),
                attractor_strength=f(constants["attractor_strength"]), attractor_x=f(attractor[0]),
                attractor_y=f(attractor[1]))


def _ws_bytes(B, S, per_env, n_steps, ckpt_every):  # pragma: no cover
    d = RolloutDesc()
    d.B, d.S, d.per_env, d.n_steps, d.ckpt_every = B, S, per_env, n_steps, ckpt_every
    return int(lib.doper_rollout_workspace_bytes(ctypes.byref(d)))


def _scratch_bytes(B, S, per_env, n_steps, ckpt_every):  # pragma: no cover
    d = RolloutDesc()
    d.B, d.S, d.per_env, d.n_steps, d.ckpt_every = B, S, per_env, n_steps, ckpt_every
    return int(lib.doper_rollout_bwd_scratch_bytes(ctypes.byref(d)))


def prepare_scene(segments, max_radius=0.0):  # pragma: no cover
    per_env = segments.ndim == 4
    n, S = (segments.shape[0], segments.shape[1]) if per_env else (1, segments.shape[0])
    nbytes = int(lib.doper_scene_workspace_bytes(n, S))
    return jax.ffi.ffi_call("doper_scene_prepare_ffi", jax.ShapeDtypeStruct((nbytes,), jnp.uint8))(segments, max_radius=np.float32(max_radius))


@partial(jax.custom_vjp, nondiff_argnums=(3, 4, 5, 6, 7, 8, 9))  # pragma: no cover
def rollout(scene_ws, x0, v0, S, per_env, n_steps, ckpt_every, dt, constants_tuple, attractor):
    return _fwd(scene_ws, x0, v0, S, per_env, n_steps, ckpt_every, dt, constants_tuple, attractor)[0]


def _fwd(scene_ws, x0, v0, S, per_env, n_steps, ckpt_every, dt, constants_tuple, attractor):  # pragma: no cover
    B = x0.shape[0]
    out = (jax.ShapeDtypeStruct((B, 2), jnp.float32), jax.ShapeDtypeStruct((B, 2), jnp.float32),
           jax.ShapeDtypeStruct((_ws_bytes(B, S, per_env, n_steps, ckpt_every),), jnp.uint8))
    xT, vT, ws = jax.ffi.ffi_call("doper_rollout_fwd_ffi", out)(
        scene_ws, x0, v0, **_attrs(S, per_env, n_steps, ckpt_every, dt, dict(constants_tuple), attractor))
    return (xT, vT), (scene_ws, ws)


def _bwd(S, per_env, n_steps, ckpt_every, dt, constants_tuple, attractor, res, cot):  # pragma: no cover
    scene_ws, ws = res
    xT_bar, vT_bar = cot
    B = xT_bar.shape[0]
    out = (jax.ShapeDtypeStruct((B, 2), jnp.float32), jax.ShapeDtypeStruct((B, 2), jnp.float32),
           jax.ShapeDtypeStruct((_scratch_bytes(B, S, per_env, n_steps, ckpt_every),), jnp.uint8))
    x0_bar, v0_bar, _scratch = jax.ffi.ffi_call("doper_rollout_bwd_ffi", out)(
        scene_ws, ws, xT_bar, vT_bar, **_attrs(S, per_env, n_steps, ckpt_every, dt, dict(constants_tuple), attractor))
    return None, x0_bar, v0_bar  # no gradient w.r.t. the scene (reference: argnums=4 only)


rollout.defvjp(_fwd, _bwd)  # pragma: no cover
