"""ORACLE (test infrastructure only): CPU restatement of the reference's per-frame loop AROUND the fusion path - the tracker's
front end and Gauss-Newton driver - assembled from the pinned oracle pieces.

  reference: pytorch/system/tracker.py
      _make_image_pyramid :41-56    -> make_image_pyramid()      (torch CPU interpolate + imgproc_oracle.gradient_xy)
      track_camera        :74-129   -> OracleTracker.track_camera (imgproc_oracle.unproject_depth, pcproc_oracle.*, synthetic.box_filter)
      compute_rgb_Hg      :131-172  -> imgproc_oracle.compute_rgb_Hg
      compute_sdf_Hg      :174-218  -> dif_oracle.compute_sdf_Hg
      gauss_newton        :220-283  -> OracleTracker.gauss_newton
  reference: pytorch/main.py:71-94 (track, then integrate with the tracked pose) -> run_loop()

Every piece it calls is pinned on its own; this file only sequences them the way the reference does, so it is
the CPU stand-in for the FULL loop: bench.py times it beside the GPU `full_loop` extra, tests compare tracked poses.
"""
from __future__ import annotations

import copy

import numpy as np
import torch

from . import dif_oracle as O
from . import imgproc_oracle as I
from . import pcproc_oracle as P


def make_image_pyramid(intensity: np.ndarray, depth: np.ndarray):
    """tracker.py:41-56 on CPU tensors: 3 levels of (bilinear intensity, nearest depth) + Sobel gradients."""
    F = torch.nn.functional
    i0 = torch.from_numpy(np.ascontiguousarray(intensity, np.float32))[None, None]
    d0 = torch.from_numpy(np.ascontiguousarray(depth, np.float32))[None, None]
    h, w = i0.shape[-2:]
    i1 = F.interpolate(i0, (h // 2, w // 2), mode="bilinear")
    d1 = F.interpolate(d0, (h // 2, w // 2), mode="nearest")
    i2 = F.interpolate(i1, (h // 4, w // 4), mode="bilinear")
    d2 = F.interpolate(d1, (h // 4, w // 4), mode="nearest")
    ints = [t[0, 0].numpy() for t in (i0, i1, i2)]
    deps = [t[0, 0].numpy() for t in (d0, d1, d2)]
    return ints, deps, [I.gradient_xy(t) for t in ints]


class OracleTracker:
    def __init__(self, omap: O.OracleMap, iter_config, sdf_robust_k=5.0, subsample=0.5,
                 rgb=dict(weight=500.0, robust_kernel=None, robust_k=0.01, min_grad_scale=0.0, max_depth_delta=0.2)):
        self.map, self.iter_config, self.sdf_robust_k, self.subsample, self.rgb = omap, iter_config, sdf_robust_k, subsample, dict(rgb)
        self.last_intensity = self.last_depth = None
        self.all_pd_pose = []                                  # [(R, t)] float64
        self.last_processed_pc = None
        self.n_sdf = self.n_rgb = 0

    # tracker.py:88-117
    def preprocess(self, depth0: np.ndarray, fx, fy, cx, cy):
        F = torch.nn.functional
        s = self.subsample
        d = F.interpolate(torch.from_numpy(np.ascontiguousarray(depth0, np.float32))[None, None], scale_factor=s, mode="nearest",
                          recompute_scale_factor=False)[0, 0].numpy()
        pc = I.unproject_depth(d, fx * s, fy * s, cx * s, cy * s)
        pc = np.concatenate([pc, np.zeros(pc.shape[:2] + (1,), np.float32)], -1).reshape(-1, 4)
        pc = pc[~np.isnan(pc[:, 0])]
        pc = pc[P.remove_radius_outlier(pc, 16, 0.05)]
        nrm = P.estimate_normals(pc, 16, 0.1, [0.0, 0.0, 0.0])
        ok = ~np.isnan(nrm[:, 0])
        from difusion_b200 import synthetic as S                # box_filter restates tracker.py:13-23 (pinned with the map fixtures)
        return S.box_filter(pc[ok, :3], nrm[ok], 0.02)

    # tracker.py:220-283
    def gauss_newton(self, init, pyr, obs_xyz, K, intr):
        R_last, t_last = self.all_pd_pose[-1]
        Ri, ti = init
        cur = (R_last.T @ Ri, R_last.T @ (ti - t_last))        # last.inv().dot(init)
        last_delta = copy.deepcopy(cur)
        ints, deps, grads = pyr
        i_iter = 0
        for group in self.iter_config:
            last_energy = np.inf
            for i_iter in list(range(group["n"])) + [-1]:
                H, g, energy = np.zeros((6, 6)), np.zeros(6), 0.0
                for loss in group["type"]:
                    if loss[0] == "sdf":
                        sH, sg, sE = O.compute_sdf_Hg(self.map, R_last, t_last, cur[0], cur[1], obs_xyz, self.sdf_robust_k, no_grad=i_iter == -1)
                        self.n_sdf += 1
                        energy += sE
                        if i_iter != -1:
                            H += sH; g += sg
                    elif loss[0] == "rgb":
                        lv = loss[1]
                        rH, rg, rE, _ = I.compute_rgb_Hg(self.last_intensity[lv], self.last_depth[lv], ints[lv], deps[lv], grads[lv], intr, K,
                                                         cur[0], cur[1], self.rgb["min_grad_scale"], self.rgb["max_depth_delta"], self.rgb["weight"],
                                                         self.rgb["robust_kernel"], self.rgb["robust_k"], no_grad=i_iter == -1)
                        self.n_rgb += 1
                        energy += rE
                        if i_iter != -1:
                            H += rH; g += rg
                    else:
                        raise NotImplementedError(loss[0])
                if energy > last_energy:
                    cur = last_delta
                    break
                last_delta = copy.deepcopy(cur)
                last_energy = energy
                if i_iter != -1:
                    Rx, tx = O.se3_exp(np.linalg.solve(H, -g))
                    cur = (Rx @ cur[0], Rx @ cur[1] + tx)       # from_twist(xi) @ cur_delta_pose
        return R_last @ cur[0], R_last @ cur[1] + t_last

    # tracker.py:74-129
    def track_camera(self, rgb: np.ndarray, depth: np.ndarray, fx, fy, cx, cy, set_pose=None):
        pyr = make_image_pyramid(rgb.mean(-1), depth)
        pc, nrm = self.preprocess(pyr[1][0], fx, fy, cx, cy)
        self.last_processed_pc = [pc, nrm]
        if set_pose is not None:
            pose = (np.asarray(set_pose[0], float), np.asarray(set_pose[1], float))
        else:
            K = np.asarray([[fx, 0.0, cx], [0.0, fy, cy], [0.0, 0.0, 1.0]])
            pose = self.gauss_newton(self.all_pd_pose[-1], pyr, pc, K, [fx, fy, cx, cy])
        self.last_intensity, self.last_depth = pyr[0], pyr[1]
        self.all_pd_pose.append(pose)
        return pose


def run_loop(weights, map_args, frames, iter_config, fx, fy, cx, cy):
    """main.py:71-94 without GUI / meshing: frames = [(rgb (H,W,3), depth (H,W), (R, t) ground truth)]; the first pose is set, the
    others are tracked; every frame is integrated with its pose.  Returns (poses, tracker, map)."""
    omap = O.OracleMap(weights, map_args)
    trk = OracleTracker(omap, iter_config)
    poses = []
    for f, (rgb, depth, gt) in enumerate(frames):
        R, t = trk.track_camera(rgb, depth, fx, fy, cx, cy, set_pose=gt if f == 0 else None)
        pc, nrm = trk.last_processed_pc
        R32, t32 = R.astype(np.float32), t.astype(np.float32)
        omap.integrate_keyframe((pc @ R32.T + t32[None]).astype(np.float32), (nrm @ R32.T).astype(np.float32))
        poses.append((R, t))
    return poses, trk, omap
