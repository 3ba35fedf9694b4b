"""Is the front end + integrate deterministic across repeated runs in one process?  (pc, normals, map state bitwise)"""
import argparse, sys
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


base = dict(sdf=dict(robust_kernel="huber", robust_k=5.0, subsample=0.5),
            rgb=dict(weight=500.0, robust_kernel=None, robust_k=0.01, min_grad_scale=0.0, max_depth_delta=0.2), iter_config=[])
R, t = S.orbit_pose(0, 200)
rgb, depth = S.render_rgbd(sc, R, t, step=1)
rgb_d, depth_d, gt = torch.from_numpy(rgb).to(dev), torch.from_numpy(depth).to(dev), Isometry(q=Rotation(matrix=R), t=t)
runs = []
for rep in range(4):
    m = DenseIndexedMap(model, sc.map_args(), 29, dev)
    trk = SDFTracker(m, argparse.Namespace(**base))
    pose = trk.track_camera(rgb_d, depth_d, Calib(), set_pose=gt)
    pc, nrm = trk.last_processed_pc
    m.integrate_keyframe(pose @ pc, pose.rotation @ nrm)
    n = m.n_occupied
    runs.append((pc.clone(), nrm.clone(), m.indexer.clone(), m.latent_vecs[:n].clone(), m.voxel_obs_count[:n].clone()))
    if rep:
        a, b = runs[0], runs[rep]
        same_shape = a[0].shape == b[0].shape
        print(f"rep {rep}: pc {tuple(b[0].shape)} same shape {same_shape}",
              "pc bitwise", same_shape and torch.equal(a[0], b[0]), "normals bitwise", same_shape and torch.equal(a[1], b[1]),
              "max |dn|", float((a[1] - b[1]).abs().max()) if same_shape else None,
              "indexer", torch.equal(a[2], b[2]), "obs", a[4].shape == b[4].shape and torch.equal(a[4], b[4]),
              "max |dlat|", float((a[3] - b[3]).abs().max()) if a[3].shape == b[3].shape else None)
