"""cProfile of the full reference loop through the mirror (track_camera + integrate + mesh) on synthetic RGB-D; GPU box."""
import argparse, cProfile, pstats, sys, time
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parent.parent; sys.path.insert(0, str(ROOT))
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


n_full = 16
imgs = []
for f in range(n_full):
    R, t = S.orbit_pose(f, 200)
    rgb, depth = S.render_rgbd(sc, R, t, step=1)
    imgs.append((torch.from_numpy(rgb).to(dev), torch.from_numpy(depth).to(dev), Isometry(q=Rotation(matrix=R), t=t)))
args = argparse.Namespace(sdf=dict(robust_kernel="huber", robust_k=5.0, subsample=0.5),
                          rgb=dict(weight=500.0, robust_kernel=None, robust_k=0.01, min_grad_scale=0.0, max_depth_delta=0.2),
                          iter_config=[{"n": 10, "type": [["rgb", 2]]}, {"n": 10, "type": [["sdf"], ["rgb", 1]]}, {"n": 50, "type": [["sdf"], ["rgb", 0]]}])


def run():
    m = DenseIndexedMap(model, sc.map_args(), 29, dev, initial_capacity=1 << 19)
    trk = SDFTracker(m, args)
    for f, (rgb_d, depth_d, gt) in enumerate(imgs):
        pose = trk.track_camera(rgb_d, depth_d, Calib(), set_pose=gt if f == 0 else None)
        pc_c, n_c = trk.last_processed_pc
        m.integrate_keyframe(pose @ pc_c, pose.rotation @ n_c)
    torch.cuda.synchronize()


run()
t0 = time.perf_counter(); run(); print(f"{1e3 * (time.perf_counter() - t0) / n_full:.2f} ms/frame (no meshing)")
pr = cProfile.Profile(); pr.enable(); run(); pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(28)
