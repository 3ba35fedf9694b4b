"""ORACLE (test infrastructure only): CPU restatement of the reference's point-cloud ops that sit in front of the fusion path.

  reference: pytorch/system/ext/pcproc/pcproc.cu
      remove_radius_outlier :98-105,172-196 -> remove_radius_outlier()
      estimate_normals      :107-170,198-220 -> estimate_normals()
  The reference answers both with an exact 16-NN search in a CUDA kd-tree (cuda_kdtree.cu, third-party tinyflann code vendored in
  the reference); here the same exact k-NN comes from scipy.spatial.cKDTree in float64, with the decisive squared distances
  re-evaluated in fp32 the way the kd-tree's CudaL2::dist compiles (fma(dz,dz,fma(dy,dy,dx*dx))).  The PCA normal is computed with
  numpy's symmetric eigensolver instead of the reference's closed-form fp32 formula, so normals are compared by angle
  (tests use 1e-3 rad), not bit for bit.

Parity status: PINNED against tests/golden/ref_ext_pcproc.npz, which tests/golden/make_golden_gpu.py produced by executing the
unmodified reference extension (oracle/_ref/pcproc) on a B200.
"""
from __future__ import annotations

import numpy as np


def _d2_f32(q, p):
    """fp32 squared distance as the kd-tree evaluates it (each fma emulated in float64, rounded once)."""
    d = (q.astype(np.float32) - p.astype(np.float32)).astype(np.float32).astype(np.float64)
    t = np.float32(d[..., 0] * d[..., 0]).astype(np.float64)
    t = np.float32(d[..., 1] * d[..., 1] + t).astype(np.float64)
    return np.float32(d[..., 2] * d[..., 2] + t)


def _knn(xyz, k):
    from scipy.spatial import cKDTree
    tree = cKDTree(xyz.astype(np.float64))
    kk = min(k, xyz.shape[0])
    _, idx = tree.query(xyz.astype(np.float64), k=kk)
    idx = idx.reshape(xyz.shape[0], kk)
    d2 = _d2_f32(xyz[:, None, :], xyz[idx])
    if kk < k:                                                    # the reference pads with (inf, -1)
        d2 = np.concatenate([d2, np.full((xyz.shape[0], k - kk), np.inf, np.float32)], 1)
        idx = np.concatenate([idx, np.full((xyz.shape[0], k - kk), -1)], 1)
    order = np.argsort(d2, axis=1, kind="stable")
    return np.take_along_axis(d2, order, 1), np.take_along_axis(idx, order, 1)


def remove_radius_outlier(pc, nb_points: int, radius: float):
    xyz = np.asarray(pc, np.float32)[:, :3]
    d2, _ = _knn(xyz, nb_points)
    return d2[:, nb_points - 1] < np.float32(radius) * np.float32(radius)


def estimate_normals(pc, max_nn: int, radius: float, cam_xyz):
    """pcproc.cu:107-170 vectorised: neighbours 1..max_nn-1 up to the first one outside the radius, mean, covariance, eigenvector of
    the smallest eigenvalue (batched LAPACK eigh), flipped towards cam_xyz; NaN rows with fewer than 5 neighbours."""
    xyz = np.asarray(pc, np.float32)[:, :3]
    n = xyz.shape[0]
    out = np.full((n, 3), np.nan, np.float32)
    if n == 0:
        return out
    d2, idx = _knn(xyz, max_nn)
    r2 = np.float32(radius) * np.float32(radius)
    inside = d2[:, 1:] < r2
    use = np.logical_and.accumulate(inside, axis=1)               # the reference stops at the first neighbour outside the radius
    cnt = use.sum(1)
    ok = cnt >= 5
    if not ok.any():
        return out
    nb = xyz[np.where(idx[:, 1:] >= 0, idx[:, 1:], 0)].astype(np.float64)          # (n, k-1, 3)
    w = use[..., None].astype(np.float64)
    mean = (nb * w).sum(1) / np.maximum(cnt, 1)[:, None]
    c = (nb - mean[:, None, :]) * w
    cov = np.einsum("nka,nkb->nab", c, c)
    _, vec = np.linalg.eigh(cov[ok])
    nrm = vec[:, :, 0]
    flip = np.einsum("na,na->n", nrm, xyz[ok].astype(np.float64) - np.asarray(cam_xyz, np.float64)) > 0
    nrm[flip] = -nrm[flip]
    out[ok] = nrm
    return out
