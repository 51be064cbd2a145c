"""Shared helpers for the parity tests."""
import numpy as np


def same_f32(a, b) -> np.ndarray:
    """Elementwise 'bit-exact' for float32: identical bits, or both NaN, or +0 == -0."""
    a = np.asarray(a, np.float32)
    b = np.asarray(b, np.float32)
    return (a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b)) | ((a == 0) & (b == 0))


def assert_bitexact(a, b, what=""):
    ok = same_f32(a, b)
    if not ok.all():
        bad = np.argwhere(~ok)
        i = tuple(bad[0])
        raise AssertionError(f"{what}: {len(bad)} of {ok.size} float32 values differ; first at {i}: "
                             f"{np.asarray(a)[i]!r} vs {np.asarray(b)[i]!r}")


def grad_rel_err(g, g_ref):
    """Per-ball error relative to the per-ball gradient norm (the gradient is a 2-vector)."""
    g = np.asarray(g, np.float64)
    g_ref = np.asarray(g_ref, np.float64)
    nrm = np.linalg.norm(g_ref, axis=-1, keepdims=True)
    return np.abs(g - g_ref).max(axis=-1) / np.maximum(nrm[..., 0], 1e-30)
