"""Synthetic RGB-D input for the DI-Fusion hot path (numpy, host side; no dataset ships with the reference).

Scenes follow SURVEY.md section 8(d):

* ``S0``  sphere r=1.2 m centred (0,0,2) seen from the origin, 32^3 PLIVox grid at 0.1 m.
* ``S1``  box room + sphere, 8x5x6 m bounds, 200-frame yaw orbit; 27k points per 640x480 frame.

The frame pipeline restates the *arithmetic* of the reference's pre-processing that decides
which points reach ``integrate_keyframe`` (it is not the kd-tree pre-processing itself, which is
out of scope): nearest 1/2 sub-sampling and pin-hole unprojection with scaled
intrinsics (reference ``system/tracker.py:88-97``, ``ext/imgproc/imgproc.cu:5-24``), depth clipping
(``main.py:67-68``) and the 2 cm box filter (``tracker.py:13-23``).  Normals are analytic.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

# ICL-NUIM intrinsics, reference dataset/production/icl_nuim.py:16
ICL_FX, ICL_FY, ICL_CX, ICL_CY = 481.2, 480.0, 319.5, 239.5
IMG_W, IMG_H = 640, 480


@dataclass
class Scene:
    name: str
    bound_min: list
    bound_max: list
    voxel_size: float
    prune_min_vox_obs: int
    ignore_count_th: float
    encoder_count_th: float = 600.0
    sphere_c: tuple = (0.0, 0.0, 2.0)
    sphere_r: float = 1.2
    room: tuple | None = None            # (xmin, xmax, ymin, ymax, zmin, zmax) or None
    depth_cut: tuple = (0.5, 5.0)
    extra: dict = field(default_factory=dict)

    def map_args(self):
        import argparse
        return argparse.Namespace(bound_min=list(self.bound_min), bound_max=list(self.bound_max),
                                  voxel_size=self.voxel_size, prune_min_vox_obs=self.prune_min_vox_obs,
                                  ignore_count_th=self.ignore_count_th, encoder_count_th=self.encoder_count_th,
                                  optim_n_iters=0)


def scene_S0() -> Scene:
    return Scene("S0", [-1.6, -1.6, 0.4], [1.6, 1.6, 3.6], 0.1, 16, 16.0)


def scene_S1(voxel_size: float = 0.05) -> Scene:
    coarse = voxel_size >= 0.1
    return Scene("S1", [-3.5, -2.5, -0.5], [4.5, 2.5, 5.5], voxel_size,
                 16 if coarse else 2, 16.0 if coarse else 4.0,
                 sphere_c=(0.3, 0.2, 2.0), sphere_r=0.6, room=(-2.0, 2.0, -1.5, 1.5, -1.0, 3.0))


def yaw_pose(yaw_rad: float, t=(0.0, 0.0, 0.0)):
    """Camera-to-world rotation (yaw about +y) and translation."""
    c, s = np.cos(yaw_rad), np.sin(yaw_rad)
    R = np.array([[c, 0.0, s], [0.0, 1.0, 0.0], [-s, 0.0, c]])
    return R, np.asarray(t, dtype=float)


def orbit_pose(frame: int, n_frames: int = 200):
    """Smooth yaw sweep -20 deg -> +20 deg over the stream (<=0.32 deg and <=1 cm per frame), SURVEY 8(d) S1."""
    ph = np.pi * (frame % (2 * n_frames)) / n_frames
    yaw = -np.deg2rad(20.0) * np.cos(ph)
    t = (0.25 * np.sin(ph), 0.05 * np.sin(2 * ph), -0.25 * np.sin(ph))
    return yaw_pose(yaw, t)


def render_depth(scene: Scene, R: np.ndarray, t: np.ndarray, noise_sigma: float = 0.0, seed: int = 0, step: int = 1):
    """Analytic ray cast of every `step`-th pixel of the 640x480 image: (H/step, W/step) float32 depth (NaN = invalid)
    and world-frame normals.  step=2 equals rendering at full resolution followed by the tracker's nearest 1/2 sub-sampling."""
    u, v = np.meshgrid(np.arange(0, IMG_W, step, dtype=np.float64), np.arange(0, IMG_H, step, dtype=np.float64))
    d_c = np.stack([(u - ICL_CX) / ICL_FX, (v - ICL_CY) / ICL_FY, np.ones_like(u)], axis=-1)   # z-depth param
    d_w = d_c @ R.T
    o = t.reshape(1, 1, 3)
    best = np.full(u.shape, np.inf)
    nrm = np.zeros(u.shape + (3,))

    # sphere
    c = np.asarray(scene.sphere_c, dtype=float)
    oc = o - c
    a = (d_w * d_w).sum(-1)
    b = 2.0 * (d_w * oc).sum(-1)
    cc = (oc * oc).sum(-1) - scene.sphere_r ** 2
    disc = b * b - 4 * a * cc
    ok = disc > 0
    s = np.where(ok, (-b - np.sqrt(np.where(ok, disc, 0.0))) / (2 * a), np.inf)
    ok &= s > 1e-6
    s = np.where(ok, s, np.inf)
    hit = o + d_w * np.where(np.isfinite(s), s, 0.0)[..., None]
    n_s = (hit - c) / scene.sphere_r
    upd = s < best
    best = np.where(upd, s, best)
    nrm = np.where(upd[..., None], n_s, nrm)

    # room (seen from inside): nearest positive exit through each wall
    if scene.room is not None:
        lo = np.array(scene.room[0::2]); hi = np.array(scene.room[1::2])
        for ax in range(3):
            for bound, sign in ((lo[ax], 1.0), (hi[ax], -1.0)):
                with np.errstate(divide="ignore", invalid="ignore"):
                    s = (bound - o[..., ax]) / d_w[..., ax]
                p = o + d_w * np.where(np.isfinite(s), s, 0.0)[..., None]
                inside = np.ones(u.shape, dtype=bool)
                for ox in range(3):
                    if ox != ax:
                        inside &= (p[..., ox] >= lo[ox] - 1e-9) & (p[..., ox] <= hi[ox] + 1e-9)
                ok = np.isfinite(s) & (s > 1e-6) & inside
                s = np.where(ok, s, np.inf)
                upd = s < best
                best = np.where(upd, s, best)
                n_w = np.zeros(3); n_w[ax] = sign
                nrm = np.where(upd[..., None], n_w.reshape(1, 1, 3), nrm)

    depth = best.copy()
    if noise_sigma > 0:
        depth = depth + np.random.default_rng(seed).normal(0.0, noise_sigma, depth.shape)
    depth = depth.astype(np.float32)
    bad = ~np.isfinite(best) | (depth < scene.depth_cut[0]) | (depth > scene.depth_cut[1])
    depth[bad] = np.nan
    return depth, nrm.astype(np.float32)


def box_filter(points: np.ndarray, normals: np.ndarray, voxel: float = 0.02):
    """Restates reference tracker.py:13-23 (point_box_filter): per-cell mean of points and normals,
    output ordered by the sorted unique cell key."""
    p32 = points.astype(np.float32)
    mn = p32.min(0, keepdims=True) - np.float32(voxel * 0.5)
    mx = p32.max(0, keepdims=True) + np.float32(voxel * 0.5)
    coord = np.floor((p32 - mn) / np.float32(voxel)).astype(np.int64)
    n = np.floor((mx - mn) / np.float32(voxel)).astype(np.int64)[0] + 16
    key = coord[:, 0] + coord[:, 1] * n[0] + coord[:, 2] * n[0] * n[1]
    _, inv = np.unique(key, return_inverse=True)
    m = int(inv.max()) + 1
    cnt = np.bincount(inv, minlength=m).astype(np.float64)
    out_p = np.stack([np.bincount(inv, weights=p32[:, k].astype(np.float64), minlength=m) / cnt for k in range(3)], 1)
    out_n = np.stack([np.bincount(inv, weights=normals[:, k].astype(np.float64), minlength=m) / cnt for k in range(3)], 1)
    return out_p.astype(np.float32), out_n.astype(np.float32)


def frame_points(scene: Scene, R: np.ndarray, t: np.ndarray, subsample: int = 2, noise_sigma: float = 0.0,
                 seed: int = 0, box: float = 0.02):
    """One 640x480 frame -> (pc_cam (N,3), normal_cam (N,3)) float32, as tracker.last_processed_pc would hold."""
    d, n_w = render_depth(scene, R, t, noise_sigma, seed, step=subsample)    # nearest, scale 0.5  (tracker.py:89-91)
    h, w = d.shape
    sc = 1.0 / subsample
    fx, fy, cx, cy = (np.float32(ICL_FX * sc), np.float32(ICL_FY * sc), np.float32(ICL_CX * sc), np.float32(ICL_CY * sc))
    uu, vv = np.meshgrid(np.arange(w, dtype=np.float32), np.arange(h, dtype=np.float32))
    x = (uu - cx) / fx * d                                   # imgproc.cu:17-19
    y = (vv - cy) / fy * d
    pc = np.stack([x, y, d], -1).reshape(-1, 3)
    n_c = (n_w.reshape(-1, 3).astype(np.float64) @ R).astype(np.float32)      # world -> camera frame
    keep = ~np.isnan(pc[:, 0])
    pc, n_c = pc[keep], n_c[keep]
    if box > 0:
        pc, n_c = box_filter(pc, n_c, box)
        n_c = n_c / np.maximum(np.linalg.norm(n_c, axis=1, keepdims=True), 1e-12)
    return pc.astype(np.float32), n_c.astype(np.float32)


def to_world(pc_cam: np.ndarray, n_cam: np.ndarray, R: np.ndarray, t: np.ndarray):
    """Isometry @ points as reference utils/motion_util.py:322-327 does it (fp32 matmul + translation)."""
    R32, t32 = R.astype(np.float32), t.astype(np.float32)
    return (pc_cam @ R32.T + t32[None, :]).astype(np.float32), (n_cam @ R32.T).astype(np.float32)


def stream_frames(scene: Scene, n_frames: int, noise_sigma: float = 0.0):
    """Yields (pc_cam, n_cam, R, t) for the orbit stream."""
    for f in range(n_frames):
        R, t = orbit_pose(f, 200)
        pc, n = frame_points(scene, R, t, noise_sigma=noise_sigma, seed=f)
        yield pc, n, R, t


def render_rgbd(scene: Scene, R: np.ndarray, t: np.ndarray, step: int = 1, noise_sigma: float = 0.0, seed: int = 0):
    """A synthetic RGB-D frame for the photometric term: depth as render_depth, colour from a smooth procedural texture that is
    a function of the WORLD hit point (so two views of the same surface agree, which is what rgb_odometry assumes).
    Returns rgb (H/step, W/step, 3) float32 in [0,1] and depth (H/step, W/step) float32 (NaN = invalid)."""
    depth, _ = render_depth(scene, R, t, noise_sigma, seed, step=step)
    u, v = np.meshgrid(np.arange(0, IMG_W, step, dtype=np.float64), np.arange(0, IMG_H, step, dtype=np.float64))
    d_c = np.stack([(u - ICL_CX) / ICL_FX, (v - ICL_CY) / ICL_FY, np.ones_like(u)], axis=-1)
    z = np.where(np.isnan(depth), 0.0, depth.astype(np.float64))
    p = t.reshape(1, 1, 3) + (d_c @ R.T) * z[..., None]
    base = 0.5 + 0.2 * np.sin(5.0 * p[..., 0]) * np.cos(4.0 * p[..., 1]) + 0.2 * np.sin(3.0 * p[..., 2] + 2.0 * p[..., 0])
    rgb = np.stack([base, 0.9 * base + 0.05 * np.cos(7.0 * p[..., 1]), 0.8 * base + 0.1 * np.sin(6.0 * p[..., 2])], -1)
    rgb = np.clip(rgb, 0.0, 1.0)
    rgb[np.isnan(depth)] = 0.0
    return rgb.astype(np.float32), depth


# ---- S3 (BASELINE configs[4]): a 50 m x 50 m height field at 5 cm PLIVoxes -> ~1 M active PLIVoxes ---------------------
# z = sum of sines (amplitude < 0.6 m, wavelengths 3-11 m), dense index 1000 x 1000 x 40 = 40 M cells.  Two generators:
#   * s3_terrain_points: the whole surface as (points, normals) batches, to BUILD the map (bulk integrate_keyframe calls);
#   * s3_view: one S1-sized frame (a 320 x 240 sample pattern over a ~4 m x 3 m footprint seen from above, 2 cm box filter
#     -> ~30 k points) at a position that moves along a Lissajous path, to STREAM against the built map.
S3_EXTENT = 50.0


# ---------------------------------------------------------------------------------------------------------------- scene S2
def scene_S2(radius: float = 3.15, voxel_size: float = 0.05) -> Scene:
    """SURVEY 8(d) scene S2 / BASELINE config 4: a sphere of `radius` metres seen from inside (~147 k PLIVoxes at 5 cm, R = 3.15)."""
    half = radius + 0.15
    return Scene("S2", [-half] * 3, [half] * 3, voxel_size, 2, 4.0)


def s2_sphere_points(radius: float = 3.15, n_pts: int = 3_000_000):
    """Fibonacci-lattice points on the sphere with inward normals (world frame), float32."""
    i = np.arange(n_pts) + 0.5
    phi = np.arccos(1 - 2 * i / n_pts)
    th = np.pi * (1 + 5 ** 0.5) * i
    d = np.stack([np.cos(th) * np.sin(phi), np.sin(th) * np.sin(phi), np.cos(phi)], 1)
    return (radius * d).astype(np.float32), (-d).astype(np.float32)


def scene_S3(voxel_size: float = 0.05, extent: float = S3_EXTENT) -> Scene:
    return Scene("S3", [0.0, 0.0, -1.0], [extent, extent, 1.0], voxel_size, 2, 4.0, room=None, extra={"extent": extent})


def s3_height(x, y):
    return (0.22 * np.sin(2 * np.pi * x / 7.0) * np.cos(2 * np.pi * y / 11.0) + 0.18 * np.sin(2 * np.pi * (x + 0.5 * y) / 5.0)
            + 0.12 * np.cos(2 * np.pi * (y - 0.3 * x) / 3.0))


def s3_normal(x, y):
    a, b, c = 2 * np.pi / 7.0, 2 * np.pi / 11.0, 2 * np.pi / 5.0
    d = 2 * np.pi / 3.0
    hx = 0.22 * a * np.cos(a * x) * np.cos(b * y) + 0.18 * c * np.cos(c * (x + 0.5 * y)) + 0.12 * d * 0.3 * np.sin(d * (y - 0.3 * x))
    hy = -0.22 * b * np.sin(a * x) * np.sin(b * y) + 0.18 * c * 0.5 * np.cos(c * (x + 0.5 * y)) - 0.12 * d * np.sin(d * (y - 0.3 * x))
    n = np.stack([-hx, -hy, np.ones_like(hx)], -1)
    return n / np.linalg.norm(n, axis=-1, keepdims=True)


def s3_terrain_points(extent: float = S3_EXTENT, spacing: float = 0.02, rows_per_batch: int = 400, seed: int = 0):
    """Yields (xyz, normal) float32 batches covering [m, extent - m]^2 on a jittered `spacing` lattice (>= 4 points per 5 cm cell)."""
    rng = np.random.default_rng(seed)
    m = 0.2
    xs = np.arange(m, extent - m, spacing)
    for r0 in range(0, xs.shape[0], rows_per_batch):
        x, y = np.meshgrid(xs[r0:r0 + rows_per_batch], xs, indexing="ij")
        x = x + rng.uniform(-0.3, 0.3, x.shape) * spacing
        y = y + rng.uniform(-0.3, 0.3, y.shape) * spacing
        p = np.stack([x, y, s3_height(x, y)], -1).reshape(-1, 3).astype(np.float32)
        yield p, s3_normal(x, y).reshape(-1, 3).astype(np.float32)


def s3_view(frame: int, extent: float = S3_EXTENT, n_frames: int = 200):
    """One frame against S3: returns (pc_cam, n_cam, R, t) like stream_frames().  The camera looks straight down from 3 m; its
    320 x 240 sample pattern covers a 4 m x 3 m footprint (1.25 cm pitch), filtered to 2 cm as the tracker does."""
    ph = 2 * np.pi * frame / n_frames
    cx, cy = extent / 2 + 0.35 * extent * np.sin(ph), extent / 2 + 0.35 * extent * np.sin(2 * ph + 0.3)
    u, v = np.meshgrid((np.arange(320) - 159.5) * 0.0125, (np.arange(240) - 119.5) * 0.0125)
    x, y = cx + u, cy + v
    pw = np.stack([x, y, s3_height(x, y)], -1).reshape(-1, 3)
    nw = s3_normal(x, y).reshape(-1, 3)
    R = np.array([[1.0, 0.0, 0.0], [0.0, -1.0, 0.0], [0.0, 0.0, -1.0]])          # camera z axis points down
    t = np.array([cx, cy, 3.0])
    pc = ((pw - t) @ R).astype(np.float32)                                        # world -> camera: R^T (p - t)
    nc = (nw @ R).astype(np.float32)
    pc, nc = box_filter(pc, nc, 0.02)
    nc = nc / np.maximum(np.linalg.norm(nc, axis=1, keepdims=True), 1e-12)
    return pc.astype(np.float32), nc.astype(np.float32), R, t


class ReproducibleNoise:
    """Stand-in for the `torch.randn(n, device=...)` of the reference's latent-optimisation sampling (map.py:486) that two
    implementations can share: the k-th request of n samples returns numpy default_rng(1000 + k).standard_normal(n) as float32."""

    def __init__(self):
        self.k = 0

    def numpy(self, n: int) -> np.ndarray:
        v = np.random.default_rng(1000 + self.k).standard_normal(int(n)).astype(np.float32)
        self.k += 1
        return v

    def torch_randn(self, n, device=None, dtype=None, **_):
        import torch
        return torch.from_numpy(self.numpy(n)).to(device=device)

    def __call__(self, n, device):
        return self.torch_randn(n, device=device)
