"""Soak test of the full loop (track_camera + integrate + incremental meshing every 10 frames) over N frames of the S1 orbit:
frame time, device memory and tracking error must stay flat.   python tools/soak.py [n_frames]"""
import argparse, sys, time
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parent.parent; sys.path.insert(0, str(ROOT))
from difusion_b200 import synthetic as S
from difusion_b200.network import utility as net_util
from difusion_b200.system.map import DenseIndexedMap
from difusion_b200.system.tracker import SDFTracker
from difusion_b200.utils.motion_util import Isometry, Rotation
n_frames = int(sys.argv[1]) if len(sys.argv) > 1 else 600
dev = torch.device("cuda:0")
model, _ = net_util.load_model(str(ROOT / "tests" / "golden" / "weights.npz"))
sc = S.scene_S1(0.05)


class Calib:
    fx, fy, cx, cy = S.ICL_FX, S.ICL_FY, S.ICL_CX, S.ICL_CY
    def to_K(self): return np.asarray([[self.fx, 0.0, self.cx], [0.0, self.fy, self.cy], [0.0, 0.0, 1.0]])


args = argparse.Namespace(sdf=dict(robust_kernel="huber", robust_k=5.0, subsample=0.5),
                          rgb=dict(weight=500.0, robust_kernel=None, robust_k=0.01, min_grad_scale=0.0, max_depth_delta=0.2),
                          iter_config=[{"n": 10, "type": [["rgb", 2]]}, {"n": 10, "type": [["sdf"], ["rgb", 1]]}, {"n": 50, "type": [["sdf"], ["rgb", 0]]}])
imgs = []
for f in range(100):                                       # 100 distinct views, swept back and forth
    R, t = S.orbit_pose(f, 200)
    rgb, depth = S.render_rgbd(sc, R, t, step=1)
    imgs.append((torch.from_numpy(rgb).to(dev), torch.from_numpy(depth).to(dev), Isometry(q=Rotation(matrix=R), t=t)))
m = DenseIndexedMap(model, sc.map_args(), 29, dev, initial_capacity=1 << 17)
trk = SDFTracker(m, args)
t_win, errs, mem = [], [], []
torch.cuda.synchronize(); w0 = time.perf_counter()
for f in range(n_frames):
    k = f % 198
    idx = k if k < 100 else 198 - k
    rgb_d, depth_d, gt = imgs[idx]
    pose = trk.track_camera(rgb_d, depth_d, Calib(), set_pose=gt if f == 0 else None)
    pc, nrm = trk.last_processed_pc
    m.integrate_keyframe(pose @ pc, pose.rotation @ nrm)
    if f % 10 == 9:
        mesh = m.extract_mesh(4, int(4e6), max_std=0.15)
    errs.append(float(np.linalg.norm(pose.t - gt.t)))
    if f % 100 == 99:
        torch.cuda.synchronize(); w1 = time.perf_counter()
        t_win.append(1e3 * (w1 - w0) / 100); mem.append(torch.cuda.memory_allocated(dev) / 2 ** 20); w0 = w1
        print(f"frames {f - 99:4d}..{f:4d}: {t_win[-1]:.2f} ms/frame, {mem[-1]:.0f} MiB allocated, n_occupied {m.n_occupied}, "
              f"max |t - t_gt| {1e3 * max(errs[-100:]):.1f} mm, mesh {mesh.n_triangles} triangles", flush=True)
assert max(t_win[1:]) < 1.5 * min(t_win[1:]) + 0.5, t_win
assert mem[-1] < mem[1] * 1.2 + 64, mem
assert max(errs) < 0.05, max(errs)
print("SOAK OK")
