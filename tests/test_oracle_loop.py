"""CPU: the full-loop oracle (oracle/loop_oracle.py: the reference's tracker front end + Gauss-Newton driver sequenced over the
pinned oracle pieces) tracks a synthetic RGB-D motion."""
import numpy as np

from conftest import GOLDEN


def test_loop_oracle_tracks_a_frame():
    from difusion_b200 import synthetic as S
    from oracle import dif_oracle as O, loop_oracle as Lp
    W = O.load_weights_npz(GOLDEN / "weights.npz")
    sc = S.scene_S1(0.05)
    frames = []
    for f in (50, 51):
        R, t = S.orbit_pose(f)
        rgb, depth = S.render_rgbd(sc, R, t, step=1)
        frames.append((rgb, depth, (R, t)))
    poses, trk, omap = Lp.run_loop(W, sc.map_args(), frames, [{"n": 10, "type": [["sdf"], ["rgb", 0]]}], S.ICL_FX, S.ICL_FY, S.ICL_CX, S.ICL_CY)
    pc, nrm = trk.last_processed_pc
    assert pc.shape == nrm.shape and pc.shape[0] > 15000 and np.isfinite(pc).all() and np.isfinite(nrm).all()
    assert trk.n_sdf >= 2 and trk.n_rgb == trk.n_sdf and omap.n_occupied > 3000
    moved = np.linalg.norm(frames[1][2][1] - frames[0][2][1])
    err = np.linalg.norm(poses[1][1] - frames[1][2][1])
    assert err < 0.8 * moved, (err, moved)                      # closer to the true pose than the previous pose was
    assert np.allclose(poses[0][0], frames[0][2][0]) and np.allclose(poses[1][0] @ poses[1][0].T, np.eye(3), atol=1e-9)
