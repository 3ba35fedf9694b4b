"""ORACLE (test infrastructure only): ctypes front-end of oracle/imgproc_oracle.c plus a numpy restatement of the torch part of
the reference's photometric term (system/tracker.py:131-172 compute_rgb_Hg, :58-71 _robust_kernel, :41-56 pyramids' gradients)."""
from __future__ import annotations

import ctypes

import numpy as np

from . import mc_oracle

_LIB = None


def _lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(str(mc_oracle.build()))
        P, I, F = ctypes.c_void_p, ctypes.c_int, ctypes.c_float
        _LIB.dif_oracle_gradient_xy.argtypes = [P, I, I, P]
        _LIB.dif_oracle_rgb_odometry.argtypes = [P, P, P, P, P, I, I, P, P, P, F, F, P, P]
        _LIB.dif_oracle_unproject_depth.argtypes = [P, I, I, F, F, F, F, P]
        for f in (_LIB.dif_oracle_gradient_xy, _LIB.dif_oracle_rgb_odometry, _LIB.dif_oracle_unproject_depth):
            f.restype = None
    return _LIB


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def gradient_xy(intensity):
    """reference ext op gradient_xy (photometric.cu:3-22,80-93): (H,W) -> (H,W,2), NaN border."""
    I = _f32(intensity)
    out = np.empty(I.shape + (2,), np.float32)
    _lib().dif_oracle_gradient_xy(I.ctypes.data, I.shape[0], I.shape[1], out.ctypes.data)
    return out


def rgb_odometry(prev_i, prev_d, cur_i, cur_d, dIdxy, intr, krkinv, kt, min_grad_scale, max_depth_delta, compute_J=True):
    """reference ext op rgb_odometry (photometric.cu:24-78,95-138) -> [f (H,W), J (H,W,6)]; J rows of rejected pixels are NaN here
    (uninitialised in the reference)."""
    a = [_f32(x) for x in (prev_i, prev_d, cur_i, cur_d, dIdxy)]
    h, w = a[2].shape
    intr, k, t = _f32(intr), _f32(krkinv).reshape(-1), _f32(kt).reshape(-1)
    f = np.empty((h, w), np.float32)
    J = np.full((h, w, 6), np.nan, np.float32) if compute_J else None
    _lib().dif_oracle_rgb_odometry(*[x.ctypes.data for x in a], h, w, intr.ctypes.data, k.ctypes.data, t.ctypes.data,
                                   float(min_grad_scale), float(max_depth_delta), f.ctypes.data, J.ctypes.data if compute_J else None)
    return [f, J] if compute_J else [f]


def unproject_depth(depth, fx, fy, cx, cy):
    """reference ext op unproject_depth (imgproc.cu:5-44): (H,W) -> (H,W,3); NaN depth -> x = NaN, y/z unspecified (set NaN here)."""
    d = _f32(depth)
    pc = np.full(d.shape + (3,), np.nan, np.float32)
    _lib().dif_oracle_unproject_depth(d.ctypes.data, d.shape[0], d.shape[1], float(fx), float(fy), float(cx), float(cy), pc.ctypes.data)
    return pc


def robust_kernel(x, kind, k):
    """tracker.py:58-71 on fp32 numpy."""
    x = x.astype(np.float32)
    if kind == "huber":
        w = np.ones_like(x)
        ax = np.abs(x)
        m = ax > np.float32(k)
        w[m] = np.float32(k) / ax[m]
        return w
    if kind == "tukey":
        w = np.zeros_like(x)
        m = np.abs(x) <= np.float32(k)
        w[m] = (1 - (x[m] / np.float32(k)) ** 2) ** 2
        return w
    raise NotImplementedError(kind)


def compute_rgb_Hg(prev_i, prev_d, cur_i, cur_d, dIdxy, intr, K, R, t, min_grad_scale, max_depth_delta, weight,
                   robust=None, robust_k=0.0, no_grad=False):
    """tracker.py:131-172.  K: 3x3 (calib.to_K()), (R, t): current delta pose.  Returns (H (6,6) f64, g (6,) f64, energy, M)."""
    KRKinv = K @ R @ np.linalg.inv(K)
    Kt = K @ t
    out = rgb_odometry(prev_i, prev_d, cur_i, cur_d, dIdxy, intr, KRKinv.flatten().tolist(), Kt.flatten().tolist(),
                       min_grad_scale, max_depth_delta, not no_grad)
    f_map = out[0]
    valid = ~np.isnan(f_map)
    f = f_map[valid].astype(np.float32)
    Wf = f
    J = JW = None
    if not no_grad:
        J = -out[1][valid]
        JW = J
    if robust is not None:
        w = robust_kernel(f, robust, robust_k)
        Wf = Wf * w
        JW = JW * w[:, None] if JW is not None else None
    es = 1.0 / Wf.shape[0] * weight
    energy = float((f.astype(np.float64) * Wf).sum() * es)
    if no_grad:
        return None, None, energy, int(Wf.shape[0])
    H = np.einsum("na,nb->ab", JW.astype(np.float64), J.astype(np.float64)) * es
    g = (J.astype(np.float64) * Wf[:, None]).sum(0) * es
    return H, g, energy, int(Wf.shape[0])
