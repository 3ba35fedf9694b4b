"""Trace of one frame's Gauss-Newton run through both drivers (dif_gauss_newton with DIF_GN_TRACE=1 and the host loop)."""
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


args = dict(sdf=dict(robust_kernel="huber", robust_k=5.0, subsample=0.5),
            rgb=dict(weight=500.0, robust_kernel=None, robust_k=0.01, min_grad_scale=0.0, max_depth_delta=0.2),
            iter_config=[{"n": 10, "type": [["rgb", 2]]}, {"n": 10, "type": [["sdf"], ["rgb", 1]]}, {"n": 50, "type": [["sdf"], ["rgb", 0]]}])
for host in (False, True):
    m = DenseIndexedMap(model, sc.map_args(), 29, dev)
    trk = SDFTracker(m, argparse.Namespace(**{k: (dict(v) if isinstance(v, dict) else v) for k, v in args.items()}))
    trk.host_loop = host
    if host:
        o_sdf, o_rgb = trk.compute_sdf_Hg, trk.compute_rgb_Hg
        def t_sdf(*a, **k):
            r = o_sdf(*a, **k); print(f"  host sdf it {a[0]} E {r[2]:.12g} t_delta {a[2].t}"); return r
        def t_rgb(*a, **k):
            r = o_rgb(*a, **k); print(f"  host rgb L{a[0]} E {r[2]:.12g}"); return r
        trk.compute_sdf_Hg, trk.compute_rgb_Hg = t_sdf, t_rgb
    for f in range(2):
        R, t = S.orbit_pose(f, 200)
        rgb, depth = S.render_rgbd(sc, R, t, step=1)
        gt = Isometry(q=Rotation(matrix=R), t=t)
        print(f"--- {'host' if host else 'native'} frame {f}", flush=True)
        pose = trk.track_camera(torch.from_numpy(rgb).to(dev), torch.from_numpy(depth).to(dev), Calib(), set_pose=gt if f == 0 else None)
        pc, nrm = trk.last_processed_pc
        m.integrate_keyframe(pose @ pc, pose.rotation @ nrm)
        print("pose t", pose.t, flush=True)
