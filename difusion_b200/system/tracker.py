"""``SDFTracker`` - host-side mirror of the reference tracker's point-to-implicit part (reference system/tracker.py:26-283).

On the hot path (SURVEY a-9): ``compute_sdf_Hg`` and the Gauss-Newton driver ``gauss_newton`` for 'sdf' terms.  The
reference evaluates the decoder through ``map.get_sdf`` + ``torch.autograd.grad`` and syncs three times per iteration
(tracker.py:191,210,215-216); here one fused kernel (dif_icp_linearize) returns H, g, energy in a single 352-byte readback.

The photometric term: ``compute_rgb_Hg`` (tracker.py:131-172) is ONE launch of dif_rgb_linearize (warp, residual,
Jacobian, robust weight, 6x6 reduction) instead of the reference's rgb_odometry kernel + boolean-mask compaction + einsum +
three host syncs; ``_make_image_pyramid`` (tracker.py:41-56) keeps torch's interpolate for the three small resamplings and uses
dif_gradient_xy for the Sobel gradients.

Frame pre-processing inside ``track_camera`` (tracker.py:88-117) goes through ``system.ext`` (unproject_depth,
remove_radius_outlier, estimate_normals) and ``point_box_filter``; callers that already hold a pre-processed cloud use
``track_points``.
"""
from __future__ import annotations

import copy
import os

import numpy as np
import torch

from .. import _lib
from ..utils.motion_util import Isometry, Rotation
from . import ext as _ext


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
        self._colored = None
        self.compacting_front_end = os.environ.get("DIF_FRONT_END", "masked") == "compacting"     # reference-shaped front end (6 host syncs)
        self.n_unstable = 0
        self._rgb_scratch = None
        self._pin = None
        # gauss_newton runs through dif_gauss_newton (device-side energy test / solve / pose update, one C call per frame) unless
        # host_loop is set: the reference-shaped Python loop below stays as the yardstick the native loop is tested against
        self.host_loop = os.environ.get("DIF_GN_HOST", "0") == "1"
        self._gn_scratch = None
        self._gn_mailbox = None
        self.n_sdf_linearisations = 0
        self.n_rgb_linearisations = 0
        self.last_gn = None

    def _read44(self, out: torch.Tensor) -> np.ndarray:
        """44 doubles device -> host through a pinned buffer and an event (cheaper than .cpu(): no pageable staging copy)."""
        if self._pin is None:
            self._pin = torch.empty(44, dtype=torch.float64).pin_memory()
            self._pin_np = self._pin.numpy()
            self._pin_ev = torch.cuda.Event()
        self._pin.copy_(out, non_blocking=True)
        self._pin_ev.record(torch.cuda.current_stream(out.device))
        self._pin_ev.synchronize()
        return self._pin_np.copy()

    def _sdf_robust_k(self) -> float:
        """tracker.py:58-71 for the sdf term, in the ABI's encoding: > 0 Huber(k), < 0 Tukey(-k), 0 no robust kernel."""
        kind, k = self.sdf_args.robust_kernel, float(self.sdf_args.robust_k or 0.0)
        if kind is None:
            return 0.0
        if kind not in ("huber", "tukey"):
            raise NotImplementedError(kind)                                 # as tracker.py:70-71
        if not k > 0.0:
            raise ValueError("robust_k must be positive")
        return k if kind == "huber" else -k

    # -------------------------------------------------------------------------------------------------
    def compute_sdf_Hg(self, n_iter: int, last_pose: Isometry, cur_delta_pose: Isometry, obs_xyz: torch.Tensor, no_grad: bool = False):
        """tracker.py:174-218.  Returns (H (6,6) float64 ndarray, g (6,) float64, energy float); (None, None, energy) if no_grad."""
        k = self._sdf_robust_k()
        out = self.map.icp_linearize(obs_xyz, last_pose.q.rotation_matrix, last_pose.t, cur_delta_pose.q.rotation_matrix,
                                     cur_delta_pose.t, huber_k=k, want_grad=not no_grad)
        self.n_sdf_linearisations += 1
        o = self._read44(out)                        # the only host sync of the iteration
        assert o[43] > 0                              # the reference asserts on an empty valid set (utility.py:84-85)
        if no_grad:
            return None, None, float(o[42])
        return o[:36].reshape(6, 6).astype(float), o[36:42].astype(float), float(o[42])

    # ------------------------------------------------------------------------------------------------- photometric term
    def _make_image_pyramid(self, intensity_img: torch.Tensor, depth_img: torch.Tensor):
        """tracker.py:41-56: three levels of (intensity bilinear, depth nearest) and their Sobel gradients."""
        F = torch.nn.functional
        d0_w, d0_h = intensity_img.size(1), intensity_img.size(0)
        d1_w, d1_h = d0_w // 2, d0_h // 2
        d2_w, d2_h = d1_w // 2, d1_h // 2
        d0_i = intensity_img.view(1, 1, d0_h, d0_w)
        d0_d = depth_img.view(1, 1, d0_h, d0_w)
        d1_i = F.interpolate(d0_i, (d1_h, d1_w), mode="bilinear")
        d1_d = F.interpolate(d0_d, (d1_h, d1_w), mode="nearest")
        d2_i = F.interpolate(d1_i, (d2_h, d2_w), mode="bilinear")
        d2_d = F.interpolate(d1_d, (d2_h, d2_w), mode="nearest")
        ints = [t.squeeze(0).squeeze(0).contiguous() for t in (d0_i, d1_i, d2_i)]
        deps = [t.squeeze(0).squeeze(0).contiguous() for t in (d0_d, d1_d, d2_d)]
        return ints, deps, [_ext.gradient_xy(t) for t in ints]

    def compute_rgb_Hg(self, pyramid_level: int, cur_delta_pose: Isometry, cur_intensity_pyramid: list, cur_depth_pyramid: list,
                       cur_dIdxy_pyramid: list, calib, no_grad: bool = False):
        """tracker.py:131-172.  Returns (H (6,6) float64, g (6,) float64, energy) or (None, None, energy) if no_grad."""
        a = self.rgb_args
        kind = {None: 0, "huber": 1, "tukey": 2}.get(a.robust_kernel, -1)
        if kind < 0:
            raise NotImplementedError(a.robust_kernel)                      # as tracker.py:70-71
        K = calib.to_K()
        KRKinv = K @ cur_delta_pose.q.rotation_matrix @ np.linalg.inv(K)
        Kt = K @ cur_delta_pose.t
        prev_i, prev_d = self.last_intensity[pyramid_level], self.last_depth[pyramid_level]
        cur_i, cur_d, cur_g = cur_intensity_pyramid[pyramid_level], cur_depth_pyramid[pyramid_level], cur_dIdxy_pyramid[pyramid_level]
        dev = cur_i.device
        L = _lib.lib()
        if self._rgb_scratch is None:
            self._rgb_scratch = torch.zeros(L.dif_rgb_scratch_bytes(), dtype=torch.uint8, device=dev)       # zero-filled once (ABI)
        out = torch.empty(44, dtype=torch.float64, device=dev)
        h, w = cur_i.shape
        _lib.check(L.dif_rgb_linearize(_lib.ptr(prev_i), _lib.ptr(prev_d), _lib.ptr(cur_i), _lib.ptr(cur_d), _lib.ptr(cur_g), h, w,
                                       _lib.host_floats([calib.fx, calib.fy, calib.cx, calib.cy]), _lib.host_floats(KRKinv.flatten().tolist()),
                                       _lib.host_floats(Kt.flatten().tolist()), float(a.min_grad_scale), float(a.max_depth_delta), kind,
                                       float(a.robust_k or 0.0), float(a.weight), int(not no_grad), self._rgb_scratch.data_ptr(),
                                       self._rgb_scratch.numel(), out.data_ptr(), _lib.stream_ptr(dev)), "dif_rgb_linearize")
        o = self._read44(out)                          # the only host sync of the term
        if not o[43] > 0:
            raise ZeroDivisionError("float division by zero")               # tracker.py:165 with an empty valid set
        if no_grad:
            return None, None, float(o[42])
        return o[:36].reshape(6, 6).astype(float), o[36:42].astype(float), float(o[42])

    def gauss_newton(self, init_pose: Isometry, cur_intensity_pyramid, cur_depth_pyramid, cur_dIdxy_pyramid, obs_xyz: torch.Tensor, calib):
        """tracker.py:220-283.  Default: the whole loop in ONE call of dif_gauss_newton (csrc/gn.cu): term kernels read the pose from
        device memory, the energy test, the 6x6 solve and the SE(3) update run on the device, the host thread only watches a pinned
        mailbox word.  `host_loop = True` runs the reference-shaped Python loop (one readback per term and iteration) instead."""
        if self.host_loop:
            return self._gauss_newton_host(init_pose, cur_intensity_pyramid, cur_depth_pyramid, cur_dIdxy_pyramid, obs_xyz, calib)
        L = _lib.lib()
        last_pose = self.all_pd_pose[-1]
        delta0 = last_pose.inv().dot(init_pose)
        p = _lib.GnProblem()
        x = None
        if obs_xyz is not None:
            x = obs_xyz.detach()
            if x.dtype != torch.float32 or not x.is_contiguous():
                x = x.float().contiguous()
            p.obs_xyz, p.n_obs = x.data_ptr(), x.size(0)
        dev = x.device if x is not None else cur_intensity_pyramid[0].device
        p.huber_k = self._sdf_robust_k()
        groups = self.args.iter_config
        if len(groups) > _lib.GN_MAX_GROUPS:
            raise ValueError(f"iter_config has more than {_lib.GN_MAX_GROUPS} groups")
        uses_rgb = False
        p.n_groups = len(groups)
        for gi, group in enumerate(groups):
            G = p.group[gi]
            G.n_iters, G.n_terms = int(group["n"]), len(group["type"])
            if not 1 <= G.n_terms <= _lib.GN_MAX_TERMS:
                raise ValueError("a group needs 1..4 loss terms")
            for k, loss_config in enumerate(group["type"]):
                if loss_config[0] == "sdf":
                    G.kind[k] = _lib.GN_TERM_SDF
                elif loss_config[0] == "rgb":
                    G.kind[k], G.level[k] = _lib.GN_TERM_RGB, int(loss_config[1])
                    uses_rgb = True
                else:
                    raise NotImplementedError(f"loss term {loss_config[0]!r} (tracker.py:254-262 'motion' is not used by the shipped configs)")
        keep = [x]
        if uses_rgb:
            a = self.rgb_args
            kind = {None: 0, "huber": 1, "tukey": 2}.get(a.robust_kernel, -1)
            if kind < 0:
                raise NotImplementedError(a.robust_kernel)
            K = np.asarray(calib.to_K(), dtype=float)
            Kinv = np.linalg.inv(K)
            p.n_levels = len(cur_intensity_pyramid)
            for lv in range(p.n_levels):
                Lv = p.level[lv]
                ts = (self.last_intensity[lv], self.last_depth[lv], cur_intensity_pyramid[lv], cur_depth_pyramid[lv], cur_dIdxy_pyramid[lv])
                keep.extend(ts)
                Lv.prev_i, Lv.prev_d, Lv.cur_i, Lv.cur_d, Lv.cur_grad = (_lib.ptr(t) for t in ts)
                Lv.h, Lv.w = cur_intensity_pyramid[lv].shape
            p.intr[:] = [calib.fx, calib.fy, calib.cx, calib.cy]
            p.K[:], p.Kinv[:] = K.flatten().tolist(), Kinv.flatten().tolist()
            p.min_grad_scale, p.max_depth_delta = float(a.min_grad_scale), float(a.max_depth_delta)
            p.rgb_robust, p.rgb_robust_k, p.rgb_weight = kind, float(a.robust_k or 0.0), float(a.weight)
        p.last_pose[:] = last_pose.q.rotation_matrix.flatten().tolist() + np.asarray(last_pose.t, float).tolist()
        p.init_delta[:] = delta0.q.rotation_matrix.flatten().tolist() + np.asarray(delta0.t, float).tolist()
        need = L.dif_gn_scratch_bytes(p.n_obs)
        if self._gn_scratch is None or self._gn_scratch.numel() < need or self._gn_scratch.device != dev:
            self._gn_scratch = torch.zeros(L.dif_gn_scratch_bytes(max(p.n_obs, 1 << 17)), dtype=torch.uint8, device=dev)   # zero-filled once (ABI)
            self._gn_mailbox = torch.zeros(32, dtype=torch.int64).pin_memory()
        res = _lib.GnResult()
        import ctypes
        view_ref = ctypes.byref(self.map._view()) if self.map is not None else None         # (a tracker without a map: rgb terms only)
        dec = self.map._prep.decoder.data_ptr() if self.map is not None else None
        _lib.check(L.dif_gauss_newton(view_ref, dec, ctypes.byref(p), self._gn_scratch.data_ptr(),
                                      self._gn_scratch.numel(), self._gn_mailbox.data_ptr(), ctypes.byref(res), _lib.stream_ptr(dev)),
                   "dif_gauss_newton")
        del keep
        self.n_sdf_linearisations += res.n_sdf
        self.n_rgb_linearisations += res.n_rgb
        self.last_gn = dict(last_iter=res.last_iter, status=res.status, iterations=res.n_iterations, energy=res.energy)
        if res.status == _lib.GN_EMPTY:
            if res.empty_term == 1:
                raise AssertionError("compute_sdf_Hg: no observed point falls into an observed PLIVox (utility.py:84-85)")
            raise ZeroDivisionError("float division by zero")               # tracker.py:165 with an empty valid set
        if res.status == _lib.GN_SINGULAR:
            raise np.linalg.LinAlgError("Singular matrix")
        d = np.asarray(res.delta[:], dtype=float)
        cur_delta_pose = Isometry(q=Rotation(matrix=d[:9].reshape(3, 3)), t=d[9:12])
        if res.last_iter >= 10:                             # tracker.py:276-281
            self.n_unstable += 1
            if self.n_unstable >= 3 and self.rgb_args is not None:
                self.rgb_args.weight = max(self.rgb_args.weight, 500.)
        return last_pose.dot(cur_delta_pose)

    def _gauss_newton_host(self, init_pose: Isometry, cur_intensity_pyramid, cur_depth_pyramid, cur_dIdxy_pyramid, obs_xyz: torch.Tensor, calib):
        """tracker.py:220-283 statement by statement (the yardstick for the native loop)."""
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
                    elif loss_config[0] == "rgb":
                        rH, rg, rE = self.compute_rgb_Hg(loss_config[1], cur_delta_pose, cur_intensity_pyramid, cur_depth_pyramid,
                                                         cur_dIdxy_pyramid, calib, i_iter == -1)
                        cur_energy += rE
                        if i_iter != -1:
                            H += rH
                            g += rg
                    else:
                        raise NotImplementedError(f"loss term {loss_config[0]!r} (tracker.py:254-262 'motion' is not used by the shipped configs)")
                if cur_energy > last_energy:
                    cur_delta_pose = last_delta_pose
                    break
                last_delta_pose = copy.deepcopy(cur_delta_pose)
                last_energy = cur_energy
                if i_iter != -1:
                    xi = np.linalg.solve(H, -g)
                    cur_delta_pose = Isometry.from_twist(xi) @ cur_delta_pose
        if i_iter >= 10:                                    # tracker.py:276-281
            self.n_unstable += 1
            if self.n_unstable >= 3 and self.rgb_args is not None:
                self.rgb_args.weight = max(self.rgb_args.weight, 500.)
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

    def _preprocess_compacting(self, cur_depth0: torch.Tensor, rgb_data: torch.Tensor, calib):
        """tracker.py:88-117 statement by statement: three boolean-mask compactions + the ops' own status checks = 6 host syncs."""
        F = torch.nn.functional
        cur_rgb = rgb_data.permute(2, 0, 1)
        pc_scale = self.sdf_args.subsample
        pc_data = F.interpolate(cur_depth0.unsqueeze(0).unsqueeze(0), scale_factor=pc_scale, mode="nearest",
                                recompute_scale_factor=False).squeeze(0).squeeze(0).contiguous()
        cur_rgb = F.interpolate(cur_rgb.unsqueeze(0), scale_factor=pc_scale, mode="bilinear", recompute_scale_factor=False).squeeze(0)
        pc_data = _ext.unproject_depth(pc_data, calib.fx * pc_scale, calib.fy * pc_scale, calib.cx * pc_scale, calib.cy * pc_scale)
        pc_data = torch.cat([pc_data, torch.zeros((pc_data.size(0), pc_data.size(1), 1), device=pc_data.device)], dim=-1).reshape(-1, 4)
        cur_rgb = cur_rgb.permute(1, 2, 0).reshape(-1, 3)
        nan_mask = ~torch.isnan(pc_data[..., 0])
        pc_data, cur_rgb = pc_data[nan_mask], cur_rgb[nan_mask]
        with torch.cuda.device(pc_data.device):
            valid = _ext.remove_radius_outlier(pc_data.contiguous(), 16, 0.05)
            pc_data, cur_rgb = pc_data[valid], cur_rgb[valid]
            normal_data = _ext.estimate_normals(pc_data.contiguous(), 16, 0.1, [0.0, 0.0, 0.0])
            normal_valid = ~torch.isnan(normal_data[..., 0])
            normal_data, cur_rgb, pc_data = normal_data[normal_valid], cur_rgb[normal_valid], pc_data[normal_valid, :3]
        self._colored = (pc_data, cur_rgb, None)
        return point_box_filter(pc_data, normal_data, 0.02)

    def _preprocess_masked(self, cur_depth0: torch.Tensor, rgb_data: torch.Tensor, calib):
        """The same front end without compaction: invalid pixels, outliers and points without a normal stay in place as NaN rows
        (the kernels treat a NaN row as "no point": the neighbour sets, the (distance, index) tie-breaks and the exact per-cell sums
        of the valid rows are those of the compacted cloud, so the result is bit-identical), the ops' overflow flags and the row
        count stay on the device, and ONE readback at the end returns them: 1 host sync instead of 6.  Returns None when a flag is
        set (the caller then runs the compacting path, which grows the cell budgets)."""
        F = torch.nn.functional
        pc_scale = self.sdf_args.subsample
        d = F.interpolate(cur_depth0.unsqueeze(0).unsqueeze(0), scale_factor=pc_scale, mode="nearest",
                          recompute_scale_factor=False).squeeze(0).squeeze(0).contiguous()
        pc3 = _ext.unproject_depth(d, calib.fx * pc_scale, calib.fy * pc_scale, calib.cx * pc_scale, calib.cy * pc_scale).reshape(-1, 3)
        nan3 = torch.full_like(pc3, float("nan"))
        status = []
        with torch.cuda.device(pc3.device):
            valid = _ext.remove_radius_outlier(pc3, 16, 0.05, status_out=status)
            pc3 = torch.where(valid.unsqueeze(-1), pc3, nan3)
            normals = _ext.estimate_normals(pc3, 16, 0.1, [0.0, 0.0, 0.0], status_out=status)
            ok = ~torch.isnan(normals[:, 0])
            pc3 = torch.where(ok.unsqueeze(-1), pc3, nan3)
            out_p, out_n, n_out = _ext.point_box_filter(pc3, normals, 0.02, deferred=True)
            flags = torch.cat([n_out] + status).cpu().tolist()          # the front end's only host sync
        if flags[0] < 0 or flags[1] or flags[2]:
            return None
        self._colored = (pc3, rgb_data, pc_scale)                       # last_colored_pcd is compacted on demand
        return out_p[:flags[0]], out_n[:flags[0]]

    @property
    def last_colored_pcd(self):
        """tracker.py:115: [points (M,3), colours (M,3)] of the cloud before the box filter (texture export); built on first access."""
        c = self._colored
        if c is None:
            return None
        if c[2] is None:
            return [c[0], c[1]]
        pc3, rgb_data, pc_scale = c
        F = torch.nn.functional
        cur_rgb = F.interpolate(rgb_data.permute(2, 0, 1).unsqueeze(0), scale_factor=pc_scale, mode="bilinear",
                                recompute_scale_factor=False).squeeze(0).permute(1, 2, 0).reshape(-1, 3)
        keep = ~torch.isnan(pc3[:, 0])
        self._colored = (pc3[keep], cur_rgb[keep], None)
        return [self._colored[0], self._colored[1]]

    @last_colored_pcd.setter
    def last_colored_pcd(self, v):
        self._colored = None if v is None else (v[0], v[1], None)

    def track_camera(self, rgb_data: torch.Tensor, depth_data: torch.Tensor, calib, set_pose: Isometry = None) -> Isometry:
        """tracker.py:74-129.  rgb (H,W,3) f32, depth (H,W) f32 with NaN = invalid, calib: fx/fy/cx/cy + to_K()."""
        cur_intensity = torch.mean(rgb_data, dim=-1)
        cur_intensity, cur_depth, cur_dIdxy = self._make_image_pyramid(cur_intensity, depth_data)
        res = None if self.compacting_front_end else self._preprocess_masked(cur_depth[0], rgb_data, calib)
        if res is None:
            res = self._preprocess_compacting(cur_depth[0], rgb_data, calib)
        pc_data, normal_data = res
        self.last_processed_pc = [pc_data, normal_data]
        if set_pose is not None:
            final_pose = set_pose
        else:
            assert len(self.all_pd_pose) > 0
            final_pose = self.gauss_newton(self.all_pd_pose[-1].dot(Isometry()), cur_intensity, cur_depth, cur_dIdxy, pc_data, calib)
        self.last_intensity = cur_intensity
        self.last_depth = cur_depth
        self.all_pd_pose.append(final_pose)
        return final_pose


def point_box_filter(points: torch.Tensor, normals: torch.Tensor, voxel_size: float):
    """tracker.py:13-23: per-cell mean of points and normals over a `voxel_size` grid, cells in ascending key order."""
    return _ext.point_box_filter(points, normals, voxel_size)
