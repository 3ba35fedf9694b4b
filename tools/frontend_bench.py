"""Timing of the SURVEY 8(f) rows around the fusion path, OURS vs THE REFERENCE'S OWN CUDA EXTENSIONS (oracle/_ref, built unmodified
from /root/reference) on the same device tensors: frame pre-processing (unproject, radius outlier, normals, box filter) and the
photometric term (gradients, rgb_odometry + reduction), 640x480 frame of scene S1, tracker sub-sampling 0.5 -> 320x240 cloud.
CUDA events, 5 warm-ups, median of 30.  Prints one JSON object.      python tools/frontend_bench.py > gpurun_out/frontend.json"""
import argparse, json, sys
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parents[1]; sys.path.insert(0, str(ROOT))
from difusion_b200 import synthetic as S
from difusion_b200.system import ext
from difusion_b200.system.tracker import SDFTracker
from difusion_b200.utils.motion_util import Isometry, Rotation
from oracle import build_ref

dev = torch.device("cuda:0")
ref = {n: build_ref.load_module(n) for n in ("imgproc", "pcproc") if build_ref.available(n)}


def timed(fn, reps=30, warm=5):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return float(np.median(ts))


sc = S.scene_S1(0.05)
(R0, t0), (R1, t1) = S.orbit_pose(30), S.orbit_pose(31)
rgb0, d0 = S.render_rgbd(sc, R0, t0, step=1, noise_sigma=0.002, seed=1)
rgb1, d1 = S.render_rgbd(sc, R1, t1, step=1, noise_sigma=0.002, seed=2)
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
i0, i1, dd0, dd1 = t(rgb0.mean(-1)), t(rgb1.mean(-1)), t(d0), t(d1)
half = torch.nn.functional.interpolate(dd1[None, None], scale_factor=0.5, mode="nearest", recompute_scale_factor=False)[0, 0].contiguous()
fx, fy, cx, cy = S.ICL_FX / 2, S.ICL_FY / 2, S.ICL_CX / 2, S.ICL_CY / 2
out = {"unit": "us per call (median of 30, CUDA events)", "cloud": None, "rows": {}}


def row(name, ours, theirs=None, note=""):
    r = {"ours_us": timed(ours)}
    if theirs is not None:
        r["reference_ext_us"] = timed(theirs)
        r["speedup"] = r["reference_ext_us"] / r["ours_us"]
    if note:
        r["note"] = note
    out["rows"][name] = r


row("unproject_depth 320x240", lambda: ext.unproject_depth(half, fx, fy, cx, cy), (lambda: ref["imgproc"].unproject_depth(half, fx, fy, cx, cy)) if "imgproc" in ref else None)
pc = ext.unproject_depth(half, fx, fy, cx, cy)
pc4 = torch.cat([pc, torch.zeros_like(pc[..., :1])], -1).reshape(-1, 4)
pc4 = pc4[~torch.isnan(pc4[:, 0])].contiguous()
out["cloud"] = int(pc4.size(0))
row("remove_radius_outlier(16, 0.05)", lambda: ext.remove_radius_outlier(pc4, 16, 0.05),
    (lambda: ref["pcproc"].remove_radius_outlier(pc4, 16, 0.05)) if "pcproc" in ref else None, "reference = thrust kd-tree build + 16-NN search per call")
kept = pc4[ext.remove_radius_outlier(pc4, 16, 0.05)].contiguous()
row("estimate_normals(16, 0.1)", lambda: ext.estimate_normals(kept, 16, 0.1, [0.0, 0.0, 0.0]),
    (lambda: ref["pcproc"].estimate_normals(kept, 16, 0.1, [0.0, 0.0, 0.0])) if "pcproc" in ref else None, "reference = second kd-tree build + search + PCA")
nrm = ext.estimate_normals(kept, 16, 0.1, [0.0, 0.0, 0.0])
okn = ~torch.isnan(nrm[:, 0])
p3, n3 = kept[okn, :3].contiguous(), nrm[okn].contiguous()


def torch_box(points, normals, vs=0.02):            # tracker.py:13-23 with index_add_ standing in for torch_scatter
    mn = torch.min(points, dim=0, keepdim=True).values - vs * 0.5
    mx = torch.max(points, dim=0, keepdim=True).values + vs * 0.5
    rc = torch.floor((points - mn) / vs).long()
    n_x, n_y, n_z = (torch.floor((mx - mn) / vs).long() + 16).cpu().numpy().tolist()[0]
    key = rc[:, 0] + rc[:, 1] * n_x + rc[:, 2] * n_x * n_y
    _, inv, cnt = torch.unique(key, return_inverse=True, return_counts=True)
    m = cnt.numel()
    return (torch.zeros(m, 3, device=dev).index_add_(0, inv, points) / cnt[:, None], torch.zeros(m, 3, device=dev).index_add_(0, inv, normals) / cnt[:, None])


row("point_box_filter(0.02)", lambda: ext.point_box_filter(p3, n3, 0.02), lambda: torch_box(p3, n3), "reference = torch ops of tracker.py:13-23 (sort-based unique + scatter)")
out["box_filter_rows"] = int(ext.point_box_filter(p3, n3, 0.02)[0].size(0))
row("gradient_xy 640x480", lambda: ext.gradient_xy(i1), (lambda: ref["imgproc"].gradient_xy(i1)) if "imgproc" in ref else None)
g1 = ext.gradient_xy(i1)
K = np.array([[S.ICL_FX, 0, S.ICL_CX], [0, S.ICL_FY, S.ICL_CY], [0, 0, 1.0]])
Rd, td = R0.T @ R1, R0.T @ (t1 - t0)
intr = [S.ICL_FX, S.ICL_FY, S.ICL_CX, S.ICL_CY]
krk, kt = (K @ Rd @ np.linalg.inv(K)).flatten().tolist(), (K @ td).flatten().tolist()


class Calib:
    fx, fy, cx, cy = intr
    def to_K(self): return K


trk = SDFTracker(None, argparse.Namespace(sdf=dict(robust_kernel="huber", robust_k=5.0, subsample=0.5),
                                          rgb=dict(weight=500.0, robust_kernel=None, robust_k=0.01, min_grad_scale=0.0, max_depth_delta=0.2), iter_config=[]))
trk.last_intensity, trk.last_depth = [i0], [dd0]
delta = Isometry(q=Rotation(matrix=Rd), t=td)


def ref_rgb_Hg():                                   # tracker.py:139-172 around the reference's own rgb_odometry kernel
    f_map, J_map = ref["imgproc"].rgb_odometry(i0, dd0, i1, dd1, g1, intr, krk, kt, 0.0, 0.2, True)
    v = ~torch.isnan(f_map)
    f = f_map[v]; J = -J_map[v]
    es = 1. / f.size(0) * 500.0
    e = (f * f).sum().item() * es
    H = torch.einsum('na,nb->nab', J, J).sum(0) * es
    g = (J * f.unsqueeze(1)).sum(0) * es
    return H.cpu().numpy(), g.cpu().numpy(), e


row("compute_rgb_Hg 640x480 (incl. host readback)", lambda: trk.compute_rgb_Hg(0, delta, [i1], [dd1], [g1], Calib()), ref_rgb_Hg if "imgproc" in ref else None,
    "ours = one fused launch + 352 B readback; reference = rgb_odometry kernel + mask compaction + einsum + 3 syncs")
print(json.dumps(out))
