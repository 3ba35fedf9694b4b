"""Golden poses of the FULL loop (track_camera + integrate, reference main.py:71-94) from the CPU loop oracle
(oracle/loop_oracle.py: the reference's tracker front end and Gauss-Newton driver sequenced over the pinned oracle pieces).

    python tests/golden/make_golden_loop.py [n_frames]     ->  tests/golden/loop_poses.npz

The stream is synthetic and deterministic (difusion_b200.synthetic: scene S1, orbit_pose(f), render_rgbd at 640x480), so the
fixture holds only the oracle's tracked poses (R, t per frame, float64), the ground-truth poses and the per-frame counts of
Gauss-Newton linearisations; tests/test_gpu_frontend.py replays the same frames through the CUDA path and bounds the pose
difference frame by frame (SURVEY section 4 "Integration").  ~4 s per frame on 16 cores.
"""
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from difusion_b200 import synthetic as S                          # noqa: E402
from oracle import dif_oracle as O, loop_oracle as Lp             # noqa: E402

ITER_CONFIG = [{"n": 10, "type": [["rgb", 2]]}, {"n": 10, "type": [["sdf"], ["rgb", 1]]}, {"n": 50, "type": [["sdf"], ["rgb", 0]]}]   # shipped


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
    W = O.load_weights_npz(ROOT / "tests" / "golden" / "weights.npz")
    sc = S.scene_S1(0.05)
    frames = []
    for f in range(n):
        R, t = S.orbit_pose(f, 200)
        rgb, depth = S.render_rgbd(sc, R, t, step=1)
        frames.append((rgb, depth, (R, t)))
    t0 = time.time()
    omap = O.OracleMap(W, sc.map_args())
    trk = Lp.OracleTracker(omap, ITER_CONFIG)
    Rs, ts, n_lin = [], [], []
    for f, (rgb, depth, gt) in enumerate(frames):
        before = trk.n_sdf
        R, t = trk.track_camera(rgb, depth, S.ICL_FX, S.ICL_FY, S.ICL_CX, S.ICL_CY, set_pose=gt if f == 0 else None)
        pc, nrm = trk.last_processed_pc
        R32, t32 = R.astype(np.float32), t.astype(np.float32)
        omap.integrate_keyframe((pc @ R32.T + t32[None]).astype(np.float32), (nrm @ R32.T).astype(np.float32))
        Rs.append(R); ts.append(t); n_lin.append(trk.n_sdf - before)
        print(f"frame {f}: |t - t_gt| = {np.linalg.norm(t - gt[1]) * 1e3:.2f} mm, sdf linearisations {n_lin[-1]}, n_occupied {omap.n_occupied}, {time.time() - t0:.0f} s", flush=True)
    np.savez_compressed(ROOT / "tests" / "golden" / "loop_poses.npz", R=np.stack(Rs), t=np.stack(ts), R_gt=np.stack([f[2][0] for f in frames]),
                        t_gt=np.stack([f[2][1] for f in frames]), n_sdf=np.asarray(n_lin), n_occupied=np.int64(omap.n_occupied))


if __name__ == "__main__":
    main()
