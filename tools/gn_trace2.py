"""Isolate the terms: energy of each term at the initial pose through the host API vs through dif_gauss_newton with n = 0 (energy pass only)."""
import argparse, os, sys
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parent.parent; sys.path.insert(0, str(ROOT))
os.environ["DIF_GN_TRACE"] = "1"
from difusion_b200 import synthetic as S
from difusion_b200.network import utility as net_util
from difusion_b200.system.map import DenseIndexedMap
from difusion_b200.system.tracker import SDFTracker
from difusion_b200.utils.motion_util import Isometry, Rotation
dev = torch.device("cuda:0")
model, _ = net_util.load_model(str(ROOT / "tests" / "golden" / "weights.npz"))
sc = S.scene_S1(0.05)


class Calib:
    fx, fy, cx, cy = S.ICL_FX, S.ICL_FY, S.ICL_CX, S.ICL_CY
    def to_K(self): return np.asarray([[self.fx, 0.0, self.cx], [0.0, self.fy, self.cy], [0.0, 0.0, 1.0]])


base = dict(sdf=dict(robust_kernel="huber", robust_k=5.0, subsample=0.5),
            rgb=dict(weight=500.0, robust_kernel=None, robust_k=0.01, min_grad_scale=0.0, max_depth_delta=0.2))
m = DenseIndexedMap(model, sc.map_args(), 29, dev)
trk = SDFTracker(m, argparse.Namespace(iter_config=[], **base))
frames = []
for f in range(2):
    R, t = S.orbit_pose(f, 200)
    rgb, depth = S.render_rgbd(sc, R, t, step=1)
    frames.append((torch.from_numpy(rgb).to(dev), torch.from_numpy(depth).to(dev), Isometry(q=Rotation(matrix=R), t=t)))
pose = trk.track_camera(frames[0][0], frames[0][1], Calib(), set_pose=frames[0][2])
pc, nrm = trk.last_processed_pc
m.integrate_keyframe(pose @ pc, pose.rotation @ nrm)
# frame 1: pre-process by hand (what track_camera does before gauss_newton)
ints, deps, grads = trk._make_image_pyramid(torch.mean(frames[1][0], dim=-1), frames[1][1])
trk.track_camera(frames[1][0], frames[1][1], Calib(), set_pose=frames[1][2])       # fills last_processed_pc for frame 1 ...
pc1 = trk.last_processed_pc[0]
trk.all_pd_pose.pop()
# ... but the photometric term compares against frame 0: rebuild its pyramids
i0, d0, _ = trk._make_image_pyramid(torch.mean(frames[0][0], dim=-1), frames[0][1])
trk.last_intensity, trk.last_depth = i0, d0
last = trk.all_pd_pose[-1]
delta = Isometry()
for cfg in ([["sdf"]], [["rgb", 1]], [["rgb", 0]], [["sdf"], ["rgb", 1]]):
    print("=== terms", cfg, flush=True)
    for t in cfg:
        if t[0] == "sdf":
            print("  host sdf E", trk.compute_sdf_Hg(-1, last, delta, pc1, True)[2], " with grad:", trk.compute_sdf_Hg(0, last, delta, pc1, False)[2])
        else:
            print(f"  host rgb L{t[1]} E", trk.compute_rgb_Hg(t[1], delta, ints, deps, grads, Calib(), True)[2], " with grad:",
                  trk.compute_rgb_Hg(t[1], delta, ints, deps, grads, Calib(), False)[2])
    trk.args.iter_config = [{"n": 1, "type": cfg}]
    trk.gauss_newton(last.dot(Isometry()), ints, deps, grads, pc1, Calib())
    sys.stderr.flush()
