"""``SDFTracker`` - host-side mirror of the reference tracker's point-to-implicit part (reference system/tracker.py:26-283).

On the hot path (SURVEY a-9): ``compute_sdf_Hg`` and the Gauss-Newton driver ``gauss_newton`` for 'sdf' terms.  The
reference evaluates the decoder through ``map.get_sdf`` + ``torch.autograd.grad`` and syncs three times per iteration
(tracker.py:191,210,215-216); here one fused kernel (dif_icp_linearize) returns H, g, energy in a single 352-byte readback.

Frame pre-processing (``track_camera``: kd-tree outlier removal / normals, tracker.py:88-117) and the photometric term
(``compute_rgb_Hg``, tracker.py:131-172) are the "next" rows of SURVEY 8(f) and are not built: ``track_camera`` therefore
takes the already pre-processed point cloud through ``track_points``.
"""
from __future__ import annotations

import copy

import numpy as np
import torch

from ..utils.motion_util import Isometry


class _Args:
    def __init__(self, d):
        self.__dict__.update(d if isinstance(d, dict) else vars(d))


class SDFTracker:
    def __init__(self, map, args):
        self.map = map
        self.args = args
        self.sdf_args = _Args(args.sdf)
        self.rgb_args = _Args(args.rgb) if getattr(args, "rgb", None) is not None else None
        self.last_intensity = None
        self.last_depth = None
        self.all_pd_pose = []
        self.last_processed_pc = None
        self.cur_gt_pose = None
        self.last_colored_pcd = None
        self.n_unstable = 0

    # -------------------------------------------------------------------------------------------------
    def compute_sdf_Hg(self, n_iter: int, last_pose: Isometry, cur_delta_pose: Isometry, obs_xyz: torch.Tensor, no_grad: bool = False):
        """tracker.py:174-218.  Returns (H (6,6) float64 ndarray, g (6,) float64, energy float); (None, None, energy) if no_grad."""
        k = self.sdf_args.robust_k if self.sdf_args.robust_kernel is not None else 0.0
        if self.sdf_args.robust_kernel not in (None, "huber"):
            raise NotImplementedError("only the huber kernel is built (fusion-lr-kt.yaml:47)")
        out = self.map.icp_linearize(obs_xyz, last_pose.q.rotation_matrix, last_pose.t, cur_delta_pose.q.rotation_matrix,
                                     cur_delta_pose.t, huber_k=k, want_grad=not no_grad)
        o = out.cpu().numpy()                        # the only host sync of the iteration
        assert o[43] > 0                              # the reference asserts on an empty valid set (utility.py:84-85)
        if no_grad:
            return None, None, float(o[42])
        return o[:36].reshape(6, 6).astype(float), o[36:42].astype(float), float(o[42])

    def compute_rgb_Hg(self, *a, **k):
        raise NotImplementedError("photometric term (tracker.py:131-172) is SURVEY 8(f)-3, not part of this build")

    def gauss_newton(self, init_pose: Isometry, cur_intensity_pyramid, cur_depth_pyramid, cur_dIdxy_pyramid, obs_xyz: torch.Tensor, calib):
        """tracker.py:220-283 for iter_config entries made of 'sdf' terms."""
        last_pose = self.all_pd_pose[-1]
        cur_delta_pose = last_pose.inv().dot(init_pose)
        last_delta_pose = copy.deepcopy(cur_delta_pose)
        i_iter = 0
        for group in self.args.iter_config:
            last_energy = np.inf
            for i_iter in list(range(group["n"])) + [-1]:
                H = np.zeros((6, 6), dtype=float)
                g = np.zeros((6,), dtype=float)
                cur_energy = 0.0
                for loss_config in group["type"]:
                    if loss_config[0] == "sdf":
                        sH, sg, sE = self.compute_sdf_Hg(i_iter, last_pose, cur_delta_pose, obs_xyz, i_iter == -1)
                        cur_energy += sE
                        if i_iter != -1:
                            H += sH
                            g += sg
                    else:
                        raise NotImplementedError(f"loss term {loss_config[0]!r} is outside the hot path (SURVEY 8f)")
                if cur_energy > last_energy:
                    cur_delta_pose = last_delta_pose
                    break
                last_delta_pose = copy.deepcopy(cur_delta_pose)
                last_energy = cur_energy
                if i_iter != -1:
                    xi = np.linalg.solve(H, -g)
                    cur_delta_pose = Isometry.from_twist(xi) @ cur_delta_pose
        if i_iter >= 10:
            self.n_unstable += 1
        return last_pose.dot(cur_delta_pose)

    def track_points(self, pc_cam: torch.Tensor, normal_cam: torch.Tensor, set_pose: Isometry = None) -> Isometry:
        """The part of track_camera after pre-processing (tracker.py:117-127): store the cloud, track or set the pose."""
        self.last_processed_pc = [pc_cam, normal_cam]
        if set_pose is not None:
            final_pose = set_pose
        else:
            assert len(self.all_pd_pose) > 0
            final_pose = self.gauss_newton(self.all_pd_pose[-1].dot(Isometry()), None, None, None, pc_cam, None)
        self.all_pd_pose.append(final_pose)
        return final_pose

    def track_camera(self, rgb_data, depth_data, calib, set_pose: Isometry = None):
        raise NotImplementedError("depth pre-processing (unproject / kd-tree outliers / PCA normals, tracker.py:88-116) is SURVEY 8(f)-1; "
                                  "feed pre-processed points to track_points()")
