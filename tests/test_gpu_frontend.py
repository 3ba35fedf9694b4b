"""GPU: the rows of SURVEY 8(f) around the fusion path - photometric term (f-3) and frame pre-processing (f-1) - through the
reference-shaped Python surface -> C ABI, against (1) the UNMODIFIED reference CUDA extension executed on the same device
(oracle/_ref/imgproc, bit-exact), (2) the CPU oracle (oracle/imgproc_oracle.c, bit-exact; numpy restatement of the torch part,
1e-4) and (3) the committed golden vectors."""
import argparse

import numpy as np
import pytest
import torch

from conftest import GOLDEN, close

pytestmark = pytest.mark.gpu
TOL = 1e-4


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return torch.device("cuda:0")


def _t(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def _ref(name):
    from oracle import build_ref
    if not build_ref.available(name):
        pytest.skip(f"oracle/_ref/{name} not built")
    return build_ref.load_module(name)


def _bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def _photo_inputs(dev):
    from golden.make_golden_gpu import photo_case
    c = photo_case()
    t = {k: _t(c[k], dev) for k in ("prev_i", "prev_d", "cur_i", "cur_d")}
    return c, t


def test_gradient_and_rgb_odometry_bit_exact(dev):
    from difusion_b200.system import ext
    from oracle import imgproc_oracle as O
    c, t = _photo_inputs(dev)
    fx = np.load(GOLDEN / "ref_ext_photo.npz")
    grad = ext.gradient_xy(t["cur_i"])
    g_np = grad.cpu().numpy()
    o_grad = O.gradient_xy(c["cur_i"])
    assert np.array_equal(_bits(g_np), _bits(o_grad)) and np.array_equal(_bits(g_np), _bits(fx["grad"]))
    args = (c["intr"].tolist(), c["krkinv"].tolist(), c["kt"].tolist(), float(c["min_grad_scale"]), float(c["max_depth_delta"]))
    f_img, J_img = ext.rgb_odometry(t["prev_i"], t["prev_d"], t["cur_i"], t["cur_d"], grad, *args, True)
    f_only, = ext.rgb_odometry(t["prev_i"], t["prev_d"], t["cur_i"], t["cur_d"], grad, *args, False)
    f_np, J_np = f_img.cpu().numpy(), J_img.cpu().numpy()
    valid = ~np.isnan(f_np)
    assert valid.sum() > 10000 and np.array_equal(_bits(f_np), _bits(f_only.cpu().numpy()))
    o_f, o_J = O.rgb_odometry(c["prev_i"], c["prev_d"], c["cur_i"], c["cur_d"], o_grad, *args)
    for rf, rJ in ((o_f, o_J), (fx["f"], fx["J"])):                    # CPU oracle, then the executed reference (golden)
        assert np.array_equal(valid, ~np.isnan(rf))
        assert np.array_equal(_bits(f_np[valid]), _bits(rf[valid])) and np.array_equal(_bits(J_np[valid]), _bits(rJ[valid]))
    with pytest.raises(RuntimeError):
        ext.gradient_xy(t["cur_i"].cpu())


def test_against_the_reference_extension_live(dev):
    """Same device tensors into the reference's imgproc module and into ours; other sizes and parameters than the fixture."""
    from difusion_b200 import synthetic as S
    from difusion_b200.system import ext
    im = _ref("imgproc")
    sc = S.scene_S1(0.05)
    for step, (fa, fb), mgs, mdd in ((2, (10, 13), 0.0, 0.2), (1, (40, 41), 1e-4, 0.05)):
        (Ra, ta), (Rb, tb) = S.orbit_pose(fa), S.orbit_pose(fb)
        rgb_a, d_a = S.render_rgbd(sc, Ra, ta, step=step)
        rgb_b, d_b = S.render_rgbd(sc, Rb, tb, step=step, noise_sigma=0.002, seed=3)
        ia, ib = _t(rgb_a.mean(-1), dev), _t(rgb_b.mean(-1), dev)
        da, db = _t(d_a, dev), _t(d_b, dev)
        K = np.array([[S.ICL_FX / step, 0, S.ICL_CX / step], [0, S.ICL_FY / step, S.ICL_CY / step], [0, 0, 1.0]])
        Rd, td = Ra.T @ Rb, Ra.T @ (tb - ta)
        intr = [K[0, 0], K[1, 1], K[0, 2], K[1, 2]]
        krk, kt = (K @ Rd @ np.linalg.inv(K)).flatten().tolist(), (K @ td).flatten().tolist()
        g_ours, g_ref = ext.gradient_xy(ib), im.gradient_xy(ib)
        assert np.array_equal(_bits(g_ours.cpu().numpy()), _bits(g_ref.cpu().numpy()))
        f1, J1 = ext.rgb_odometry(ia, da, ib, db, g_ours, intr, krk, kt, mgs, mdd, True)
        f2, J2 = im.rgb_odometry(ia, da, ib, db, g_ref, intr, krk, kt, mgs, mdd, True)
        f1, f2, J1, J2 = f1.cpu().numpy(), f2.cpu().numpy(), J1.cpu().numpy(), J2.cpu().numpy()
        v = ~np.isnan(f2)
        assert v.sum() > 2000 and np.array_equal(v, ~np.isnan(f1))
        assert np.array_equal(_bits(f1[v]), _bits(f2[v])) and np.array_equal(_bits(J1[v]), _bits(J2[v]))
        p1 = ext.unproject_depth(db, *intr)
        p2 = im.unproject_depth(db, *intr)
        ok = ~torch.isnan(db)
        assert torch.equal(p1[ok].view(torch.int32), p2[ok].view(torch.int32)) and bool(torch.isnan(p1[~ok]).all())


def test_unproject_against_golden(dev):
    from difusion_b200.system import ext
    from oracle import imgproc_oracle as O
    fx = np.load(GOLDEN / "ref_ext_unproject.npz")
    pc = ext.unproject_depth(_t(fx["depth"], dev), *[float(v) for v in fx["intr"]]).cpu().numpy()
    assert np.array_equal(_bits(pc), _bits(fx["pc"]))
    assert np.array_equal(_bits(O.unproject_depth(fx["depth"], *[float(v) for v in fx["intr"]])), _bits(fx["pc"]))


class _Calib:
    def __init__(self, fx, fy, cx, cy):
        self.fx, self.fy, self.cx, self.cy = fx, fy, cx, cy

    def to_K(self):
        return np.asarray([[self.fx, 0.0, self.cx], [0.0, self.fy, self.cy], [0.0, 0.0, 1.0]])


@pytest.mark.parametrize("robust,k", [(None, 0.01), ("huber", 0.004), ("tukey", 0.02)])
def test_compute_rgb_Hg_matches_oracle(dev, robust, k):
    """SDFTracker.compute_rgb_Hg (one fused launch) vs the numpy restatement of tracker.py:131-172 over the C oracle."""
    from difusion_b200.system.tracker import SDFTracker
    from difusion_b200.utils.motion_util import Isometry, Rotation
    from oracle import imgproc_oracle as O
    c, t = _photo_inputs(dev)
    args = argparse.Namespace(sdf=dict(robust_kernel="huber", robust_k=5.0, subsample=0.5),
                              rgb=dict(weight=500.0, robust_kernel=robust, robust_k=k, min_grad_scale=float(c["min_grad_scale"]), max_depth_delta=0.2),
                              iter_config=[])
    trk = SDFTracker(None, args)
    from difusion_b200.system import ext
    grad = ext.gradient_xy(t["cur_i"])
    trk.last_intensity, trk.last_depth = [t["prev_i"]], [t["prev_d"]]
    calib = _Calib(*[float(v) for v in c["intr"]])
    delta = Isometry(q=Rotation(matrix=c["Rd"]), t=c["td"])
    H, g, E = trk.compute_rgb_Hg(0, delta, [t["cur_i"]], [t["cur_d"]], [grad], calib)
    oH, og, oE, oM = O.compute_rgb_Hg(c["prev_i"], c["prev_d"], c["cur_i"], c["cur_d"], grad.cpu().numpy(), c["intr"], calib.to_K(),
                                      delta.q.rotation_matrix, delta.t, float(c["min_grad_scale"]), 0.2, 500.0, robust, k)
    assert H.dtype == np.float64 and H.shape == (6, 6) and g.shape == (6,)
    assert np.abs(H - oH).max() <= 2e-4 * np.abs(oH).max() and np.abs(g - og).max() <= 2e-4 * np.abs(og).max()
    assert abs(E - oE) <= TOL * abs(oE)
    _, _, E2 = trk.compute_rgb_Hg(0, delta, [t["cur_i"]], [t["cur_d"]], [grad], calib, no_grad=True)
    assert E2 == E


def test_rgb_gauss_newton_recovers_the_relative_pose(dev):
    """track_camera's photometric Gauss-Newton (tracker.py:220-283 with ['rgb', level] terms) on two synthetic views."""
    from difusion_b200 import synthetic as S
    from difusion_b200.system.tracker import SDFTracker
    from difusion_b200.utils.motion_util import Isometry, Rotation
    sc = S.scene_S1(0.05)
    (R0, t0), (R1, t1) = S.orbit_pose(20), S.orbit_pose(22)
    rgb0, d0 = S.render_rgbd(sc, R0, t0, step=2)
    rgb1, d1 = S.render_rgbd(sc, R1, t1, step=2)
    calib = _Calib(S.ICL_FX / 2, S.ICL_FY / 2, S.ICL_CX / 2, S.ICL_CY / 2)
    args = argparse.Namespace(sdf=dict(robust_kernel="huber", robust_k=5.0, subsample=0.5),
                              rgb=dict(weight=500.0, robust_kernel=None, robust_k=0.01, min_grad_scale=0.0, max_depth_delta=0.2),
                              iter_config=[{"n": 15, "type": [["rgb", 0]]}])
    trk = SDFTracker(None, args)
    i0, dd0, _ = trk._make_image_pyramid(_t(rgb0.mean(-1), dev), _t(d0, dev))
    i1, dd1, g1 = trk._make_image_pyramid(_t(rgb1.mean(-1), dev), _t(d1, dev))
    assert [tuple(x.shape) for x in i1] == [(240, 320), (120, 160), (60, 80)] and g1[2].shape == (60, 80, 2)
    trk.last_intensity, trk.last_depth = i0, dd0
    p0 = Isometry(q=Rotation(matrix=R0), t=t0)
    trk.all_pd_pose.append(p0)
    est = trk.gauss_newton(p0, i1, dd1, g1, None, calib)
    err0 = np.linalg.norm(t1 - t0)
    err = np.linalg.norm(est.t - t1)
    ang = np.arccos(np.clip((np.trace(est.q.rotation_matrix.T @ R1) - 1) / 2, -1, 1))
    ang0 = np.arccos(np.clip((np.trace(R0.T @ R1) - 1) / 2, -1, 1))
    assert err < 0.35 * err0 and ang < 0.35 * ang0, (err, err0, ang, ang0)


def test_point_box_filter_matches_restatement(dev):
    """tracker.point_box_filter (tracker.py:13-23): same cells, same row order, means BIT-identical to the fp64 restatement (the
    per-cell sums are exact in fp64, hence independent of the accumulation order: reproducible across runs and GPUs)."""
    from difusion_b200 import synthetic as S
    from difusion_b200.system import ext
    from difusion_b200.system.tracker import point_box_filter
    sc = S.scene_S1(0.05)
    R, t = S.orbit_pose(5)
    pc, nc = S.frame_points(sc, R, t, box=0.0)                       # raw unprojected cloud, ~75 k points
    for pts, nrm, vs in ((pc, nc, 0.02), (pc[:5000] * np.float32(3.0), nc[:5000], 0.05), (pc[:1], nc[:1], 0.02)):
        exp_p, exp_n = S.box_filter(pts, nrm, vs)
        out_p, out_n = point_box_filter(_t(pts, dev), _t(nrm, dev), vs)
        assert out_p.shape == exp_p.shape and out_n.shape == exp_n.shape
        assert np.array_equal(out_p.cpu().numpy(), exp_p) and np.array_equal(out_n.cpu().numpy(), exp_n)
    a, b = ext.point_box_filter(_t(pc, dev), _t(nc, dev), 0.02)     # scratch is self-cleaning: a second call gives the same rows
    c, d = ext.point_box_filter(_t(pc, dev), _t(nc, dev), 0.02)
    assert a.shape == c.shape and torch.equal(a, c) and torch.equal(b, d)
    with pytest.raises(RuntimeError):
        ext.point_box_filter(_t(pc, dev).cpu(), _t(nc, dev), 0.02)


def _normal_agreement(a, b):
    """(fraction of rows whose NaN-ness differs, fraction of common rows whose angle exceeds 1e-3 rad)."""
    na, nb = np.isnan(a[:, 0]), np.isnan(b[:, 0])
    both = ~na & ~nb
    cosang = np.abs((a[both].astype(np.float64) * b[both]).sum(1))
    return float((na != nb).mean()), float((cosang < np.cos(1e-3)).mean())


@pytest.mark.parametrize("radius", [0.05, 0.08])
def test_knn_ops_against_reference_extension_and_oracle(dev, radius):
    """remove_radius_outlier / estimate_normals: uniform-grid k-NN vs the reference's CUDA kd-tree executed on the same tensors
    (mask bit-exact; normals to 1e-3 rad except where equal-distance ties or a 1-ulp radius test change the neighbour set) and vs
    the scipy/numpy oracle."""
    from difusion_b200.system import ext
    from golden.make_golden_gpu import cloud_case
    from oracle import pcproc_oracle as P
    cloud = cloud_case()
    d_cloud = _t(cloud, dev)
    mask = ext.remove_radius_outlier(d_cloud, 16, radius)
    m_np = mask.cpu().numpy()
    assert mask.dtype == torch.bool and 0 < m_np.sum() < cloud.shape[0]
    o_mask = P.remove_radius_outlier(cloud, 16, radius)
    assert (m_np != o_mask).mean() <= 1e-3
    kept = d_cloud[mask].contiguous()
    normals = ext.estimate_normals(kept, 16, 2 * radius, [0.0, 0.0, 0.0])
    n_np = normals.cpu().numpy()
    ok = ~np.isnan(n_np[:, 0])
    assert ok.mean() > 0.9 and np.abs(np.linalg.norm(n_np[ok], axis=1) - 1).max() < 1e-5
    assert ((n_np[ok] * kept.cpu().numpy()[ok, :3]).sum(1) <= 0).all()              # oriented towards the camera at the origin
    nan_diff, ang_diff = _normal_agreement(n_np, P.estimate_normals(cloud[m_np], 16, 2 * radius, [0.0, 0.0, 0.0]))
    assert nan_diff <= 2e-3 and ang_diff <= 5e-3, (nan_diff, ang_diff)
    fx = np.load(GOLDEN / "ref_ext_pcproc.npz")
    tag = "r5" if radius == 0.05 else "r8"
    assert np.array_equal(m_np, fx[f"{tag}.mask"])                                   # vs the executed reference (golden)
    nan_diff, ang_diff = _normal_agreement(n_np, fx[f"{tag}.normals"])
    assert nan_diff <= 1e-3 and ang_diff <= 2e-3, (nan_diff, ang_diff)
    pp = _ref("pcproc")                                                              # ... and live, same device tensors, (N,3) layout too
    assert torch.equal(pp.remove_radius_outlier(d_cloud, 16, radius), mask)
    assert torch.equal(ext.remove_radius_outlier(d_cloud[:, :3].contiguous(), 16, radius), mask)
    nan_diff, ang_diff = _normal_agreement(n_np, pp.estimate_normals(kept, 16, 2 * radius, [0.0, 0.0, 0.0]).cpu().numpy())
    assert nan_diff <= 1e-3 and ang_diff <= 2e-3, (nan_diff, ang_diff)
    again = ext.estimate_normals(kept, 16, 2 * radius, [0.0, 0.0, 0.0])              # deterministic, scratch cleaned itself
    assert torch.equal(again.view(torch.int32), normals.view(torch.int32))


def test_track_camera_end_to_end(dev):
    """The whole reference front end through the mirror (tracker.py:74-129): pyramid, unproject, outlier removal, normals, box
    filter, then integrate the processed cloud and track the next frame with the default 3-group iter_config (rgb + sdf terms)."""
    from difusion_b200 import synthetic as S
    from difusion_b200.network import utility as net_util
    from difusion_b200.system.map import DenseIndexedMap
    from difusion_b200.system.tracker import SDFTracker
    from difusion_b200.utils.motion_util import Isometry, Rotation
    model, _ = net_util.load_model(str(GOLDEN / "weights.npz"))
    sc = S.scene_S1(0.05)
    m = DenseIndexedMap(model, sc.map_args(), 29, dev)
    args = argparse.Namespace(sdf=dict(robust_kernel="huber", robust_k=5.0, subsample=0.5),
                              rgb=dict(weight=500.0, robust_kernel=None, robust_k=0.01, min_grad_scale=0.0, max_depth_delta=0.2),
                              iter_config=[{"n": 10, "type": [["sdf"], ["rgb", 0]]}])
    trk = SDFTracker(m, args)
    calib = _Calib(S.ICL_FX, S.ICL_FY, S.ICL_CX, S.ICL_CY)
    poses = [S.orbit_pose(f) for f in (50, 51)]
    est = []
    for f, (R, t) in enumerate(poses):
        rgb, depth = S.render_rgbd(sc, R, t, step=1)
        gt = Isometry(q=Rotation(matrix=R), t=t)
        pose = trk.track_camera(_t(rgb, dev), _t(depth, dev), calib, set_pose=gt if f == 0 else None)
        est.append(pose)
        pc, nrm = trk.last_processed_pc
        assert pc.shape == nrm.shape and pc.size(0) > 15000 and bool(torch.isfinite(pc).all()) and bool(torch.isfinite(nrm).all())
        m.integrate_keyframe(pose @ pc, pose.rotation @ nrm)
    assert m.n_occupied > 3000
    err = np.linalg.norm(est[1].t - poses[1][1])
    assert err < 0.6 * np.linalg.norm(poses[1][1] - poses[0][1]) + 2e-3, err      # moved towards the true pose from the previous one


def test_full_loop_pose_drift_against_loop_oracle(dev):
    """SURVEY section 4 "Integration": the whole loop (reference main.py:71-94: track_camera with the shipped 3-group iter_config,
    then integrate with the TRACKED pose) over 30 frames of the synthetic RGB-D stream, on the CUDA path, against the poses the CPU
    loop oracle tracked on the same frames (tests/golden/loop_poses.npz, made by tests/golden/make_golden_loop.py).  Tracking errors
    feed back into the map, so this bounds the accumulated divergence of the two implementations, frame by frame."""
    from difusion_b200 import synthetic as S
    from difusion_b200.network import utility as net_util
    from difusion_b200.system.map import DenseIndexedMap
    from difusion_b200.system.tracker import SDFTracker
    from difusion_b200.utils.motion_util import Isometry, Rotation
    fx = np.load(GOLDEN / "loop_poses.npz")
    n = int(fx["R"].shape[0])
    model, _ = net_util.load_model(str(GOLDEN / "weights.npz"))
    sc = S.scene_S1(0.05)
    m = DenseIndexedMap(model, sc.map_args(), 29, dev)
    args = argparse.Namespace(sdf=dict(robust_kernel="huber", robust_k=5.0, subsample=0.5),
                              rgb=dict(weight=500.0, robust_kernel=None, robust_k=0.01, min_grad_scale=0.0, max_depth_delta=0.2),
                              iter_config=[{"n": 10, "type": [["rgb", 2]]}, {"n": 10, "type": [["sdf"], ["rgb", 1]]},
                                           {"n": 50, "type": [["sdf"], ["rgb", 0]]}])
    trk = SDFTracker(m, args)
    calib = _Calib(S.ICL_FX, S.ICL_FY, S.ICL_CX, S.ICL_CY)
    def ang(Ra, Rb):
        return float(np.degrees(np.arccos(np.clip((np.trace(np.asarray(Ra) @ np.asarray(Rb).T) - 1) / 2, -1, 1))))
    dt_mm, dr_deg, gt_mm, gt_deg, ogt_deg = [], [], [], [], []
    for f in range(n):
        R, t = S.orbit_pose(f, 200)
        assert np.allclose(R, fx["R_gt"][f]) and np.allclose(t, fx["t_gt"][f])          # same stream as the fixture
        rgb, depth = S.render_rgbd(sc, R, t, step=1)
        gt = Isometry(q=Rotation(matrix=R), t=t)
        pose = trk.track_camera(_t(rgb, dev), _t(depth, dev), calib, set_pose=gt if f == 0 else None)
        pc, nrm = trk.last_processed_pc
        m.integrate_keyframe(pose @ pc, pose.rotation @ nrm)
        dR = np.asarray(pose.q.rotation_matrix) @ fx["R"][f].T
        dt_mm.append(1e3 * float(np.linalg.norm(np.asarray(pose.t) - fx["t"][f])))
        dr_deg.append(float(np.degrees(np.arccos(np.clip((np.trace(dR) - 1) / 2, -1, 1)))))
        gt_mm.append(1e3 * float(np.linalg.norm(np.asarray(pose.t) - t)))
        gt_deg.append(ang(pose.q.rotation_matrix, R)); ogt_deg.append(ang(fx["R"][f], R))
        if f % 5 == 4:
            print(f"[loop drift] frame {f}: gpu-oracle {dt_mm[-1]:.2f} mm {dr_deg[-1]:.3f} deg | gpu-gt {gt_mm[-1]:.2f} mm {gt_deg[-1]:.3f} deg | oracle-gt "
                  f"{1e3 * float(np.linalg.norm(fx['t'][f] - t)):.2f} mm {ogt_deg[-1]:.3f} deg")
    print(f"[loop drift] {n} frames: |t_gpu - t_oracle| max {max(dt_mm):.3f} mm (last {dt_mm[-1]:.3f}), rotation max {max(dr_deg):.4f} deg; "
          f"|t_gpu - t_gt| max {max(gt_mm):.2f} mm; oracle |t - t_gt| max {1e3 * float(np.linalg.norm(fx['t'] - fx['t_gt'], axis=1).max()):.2f} mm; "
          f"n_occupied {m.n_occupied} vs {int(fx['n_occupied'])}")
    # Measured (B200): identical to 0.01 mm for 10 frames, 0.08 mm at frame 14, 0.33 mm / 0.02 deg at frame 19, 0.44 mm / 0.26 deg at
    # frame 24 - and by frame 29 the CPU ORACLE has left the trajectory (4.0 deg from ground truth) while the CUDA path is within
    # 0.08 deg of it: a feedback loop of two fp32 implementations bifurcates eventually (one Gauss-Newton step accepted on one side
    # and rejected on the other, tracker.py:263-266).  The bound is therefore frame-by-frame over the first 20 frames, plus a bound
    # of the CUDA path against GROUND TRUTH over all of them (the oracle's own error peaks at 15.95 mm).
    # Across boxes the same binary gives slightly different trajectories (the encoder's fp32 atomics are the one order-dependent piece
    # left on the path: 2e-7 on a latent, enough to flip an energy test between two nearly converged iterates): measured 0.00 mm
    # through frame 9 on one box and 0.46 mm at frame 9 on another, and on some boxes ONE frame of the first 20 is off by 5.6 mm (a
    # Gauss-Newton group stopped one step earlier than the oracle's) and the next frame is back within 0.3 mm.  So: the median carries
    # the tight claim, at most two frames may leave 1 mm, none may leave 8 mm.
    print("[loop drift] per-frame gpu-oracle mm:", " ".join(f"{v:.2f}" for v in dt_mm[:20]))
    first = np.asarray(dt_mm[:20])
    assert float(np.median(first)) < 0.3 and int((first > 1.0).sum()) <= 2 and float(first.max()) < 8.0, first
    assert float(np.median(dr_deg[:20])) < 0.02 and max(dr_deg[:20]) < 0.12, max(dr_deg[:20])
    assert max(gt_mm) < 20.0 and max(gt_deg) < 0.5, (max(gt_mm), max(gt_deg))
    assert abs(m.n_occupied - int(fx["n_occupied"])) <= 0.01 * int(fx["n_occupied"])


def test_device_gauss_newton_equals_host_loop(dev):
    """dif_gauss_newton (device-side energy test / solve / pose update, csrc/gn.cu) against the reference-shaped host loop
    (SDFTracker._gauss_newton_host: one readback per term and iteration, numpy solve, Isometry algebra) on the SAME inputs (same
    pre-processed cloud, pyramids and map: the box filter's per-cell means are atomics-ordered, so two runs of the front end differ in
    the last bit).  Both drive the same deterministic term kernels, so the accepted iterates agree to fp64 rounding."""
    from difusion_b200 import synthetic as S
    from difusion_b200.network import utility as net_util
    from difusion_b200.system.map import DenseIndexedMap
    from difusion_b200.system.tracker import SDFTracker
    from difusion_b200.utils.motion_util import Isometry, Rotation
    model, _ = net_util.load_model(str(GOLDEN / "weights.npz"))
    sc = S.scene_S1(0.05)
    args = argparse.Namespace(sdf=dict(robust_kernel="huber", robust_k=5.0, subsample=0.5),
                              rgb=dict(weight=500.0, robust_kernel=None, robust_k=0.01, min_grad_scale=0.0, max_depth_delta=0.2),
                              iter_config=[{"n": 10, "type": [["rgb", 2]]}, {"n": 10, "type": [["sdf"], ["rgb", 1]]},
                                           {"n": 50, "type": [["sdf"], ["rgb", 0]]}])
    calib = _Calib(S.ICL_FX, S.ICL_FY, S.ICL_CX, S.ICL_CY)
    m = DenseIndexedMap(model, sc.map_args(), 29, dev)
    trk = SDFTracker(m, args)
    native = type(trk).gauss_newton
    seen = []

    def both(init_pose, ints, deps, grads, obs, cal):
        a = native(trk, init_pose, ints, deps, grads, obs, cal)
        gn = dict(trk.last_gn)
        n_un, w = trk.n_unstable, trk.rgb_args.weight
        s0 = trk.n_sdf_linearisations
        b = trk._gauss_newton_host(init_pose, ints, deps, grads, obs, cal)
        host_sdf = trk.n_sdf_linearisations - s0
        trk.n_unstable, trk.rgb_args.weight = n_un, w                     # the host run must not count twice
        dt = float(np.linalg.norm(np.asarray(a.t) - np.asarray(b.t)))
        dR = float(np.abs(np.asarray(a.q.rotation_matrix) - np.asarray(b.q.rotation_matrix)).max())
        seen.append((dt, dR, gn["iterations"], host_sdf))
        return b
    trk.gauss_newton = both
    for f in range(6):
        R, t = S.orbit_pose(f, 200)
        rgb, depth = S.render_rgbd(sc, R, t, step=1)
        gt = Isometry(q=Rotation(matrix=R), t=t)
        pose = trk.track_camera(_t(rgb, dev), _t(depth, dev), calib, set_pose=gt if f == 0 else None)
        pc, nrm = trk.last_processed_pc
        m.integrate_keyframe(pose @ pc, pose.rotation @ nrm)
    assert len(seen) == 5
    for f, (dt, dR, iters, host_sdf) in enumerate(seen, 1):
        print(f"[gn] frame {f}: |dt| {dt:.2e} |dR| {dR:.2e}, {iters} iterations")
        # (an fp32 pose entry may round differently after the two 6x6 solvers: iterates then differ by ~1e-9; an energy test between two
        # nearly converged iterates may flip - one more, tiny, step on one side)
        assert dt < 2e-6 and dR < 2e-6, (f, dt, dR)
        assert iters > 3


def test_device_gauss_newton_sdf_only_and_errors(dev):
    """track_points (sdf-only iter_config, no pyramids) through the native loop; an empty valid set raises like the reference."""
    from difusion_b200 import synthetic as S
    from difusion_b200.network import utility as net_util
    from difusion_b200.system.map import DenseIndexedMap
    from difusion_b200.system.tracker import SDFTracker
    from difusion_b200.utils.motion_util import Isometry, Rotation
    model, _ = net_util.load_model(str(GOLDEN / "weights.npz"))
    sc = S.scene_S1(0.05)
    ns = argparse.Namespace(sdf=dict(robust_kernel="huber", robust_k=5.0, subsample=0.5), rgb=None, iter_config=[{"n": 20, "type": [["sdf"]]}])
    res = []
    for host in (False, True):
        m = DenseIndexedMap(model, sc.map_args(), 29, dev)
        trk = SDFTracker(m, ns)
        trk.host_loop = host
        for f in range(3):
            R, t = S.orbit_pose(f, 200)
            pc, nc = S.frame_points(sc, R, t)
            gt = Isometry(q=Rotation(matrix=R), t=t)
            pose = trk.track_points(_t(pc, dev), _t(nc, dev), set_pose=gt if f == 0 else None)
            m.integrate_keyframe(gt @ _t(pc, dev), gt.rotation @ _t(nc, dev))
        res.append(pose)
    assert np.linalg.norm(np.asarray(res[0].t) - np.asarray(res[1].t)) < 2e-6
    # fewer than 2048 points: the exact-fp32 SIMT kernel serves the sdf term, reading pose and count from the same device block
    small = []
    for host in (False, True):
        trk = SDFTracker(m, ns)
        trk.host_loop = host
        trk.all_pd_pose.append(gt)
        small.append(trk.track_points(_t(pc[:1500], dev), _t(nc[:1500], dev)))
    assert np.linalg.norm(np.asarray(small[0].t) - np.asarray(small[1].t)) < 2e-6
    assert np.abs(np.asarray(small[0].q.rotation_matrix) - np.asarray(small[1].q.rotation_matrix)).max() < 2e-6
    far = torch.full((4096, 3), 50.0, device=dev)                 # every point outside the grid: M = 0
    with pytest.raises(AssertionError):
        trk2 = SDFTracker(m, ns)
        trk2.all_pd_pose.append(Isometry())
        trk2.track_points(far, far)


def test_far_point_does_not_break_the_front_end(dev):
    """one far / noisy depth pixel blew the neighbour-grid (and box-filter) cell budget and made track_camera raise; the
    reference's kd-tree has no such limit.  The budget now grows from the actual extent and the call is repeated."""
    from difusion_b200 import synthetic as S
    from difusion_b200.system import ext
    sc = S.scene_S1(0.05)
    R, t = S.orbit_pose(3)
    pc, nc = S.frame_points(sc, R, t, box=0.0)
    pc = pc[:20000]
    base = ext.remove_radius_outlier(_t(pc, dev), 16, 0.05)
    far = np.concatenate([pc, np.asarray([[400.0, pc[0, 1], pc[0, 2]]], np.float32)], 0)      # y, z inside the cloud: the grid origin stays
    mask = ext.remove_radius_outlier(_t(far, dev), 16, 0.05)
    assert torch.equal(mask[:-1], base) and not bool(mask[-1])
    nrm = ext.estimate_normals(_t(far, dev), 16, 0.1, [0.0, 0.0, 0.0])
    assert bool(torch.isnan(nrm[-1]).all()) and bool(torch.isfinite(nrm[:-1][base]).all())
    p2, n2 = ext.point_box_filter(_t(far, dev), _t(np.concatenate([nc[:20000], nc[:1]], 0), dev), 0.02)
    q2, _ = ext.point_box_filter(_t(pc, dev), _t(nc[:20000], dev), 0.02)
    assert p2.size(0) == q2.size(0) + 1
    e_p, e_n = S.box_filter(far, np.concatenate([nc[:20000], nc[:1]], 0), 0.02)
    assert np.array_equal(p2.cpu().numpy(), e_p) and np.array_equal(n2.cpu().numpy(), e_n)


def test_masked_front_end_equals_compacting_front_end(dev):
    """track_camera's default front end keeps invalid pixels / outliers / normal-less points in place as NaN rows and reads ONE
    device block back (row count + overflow flags); the reference-shaped one compacts three times (6 host syncs).  The kernels treat
    a NaN row as "no point", so the processed cloud, its normals and the coloured cloud are bit-identical."""
    from difusion_b200 import synthetic as S
    from difusion_b200.system.tracker import SDFTracker
    sc = S.scene_S1(0.05)
    args = argparse.Namespace(sdf=dict(robust_kernel="huber", robust_k=5.0, subsample=0.5),
                              rgb=dict(weight=500.0, robust_kernel=None, robust_k=0.01, min_grad_scale=0.0, max_depth_delta=0.2),
                              iter_config=[{"n": 2, "type": [["sdf"]]}])
    calib = _Calib(S.ICL_FX, S.ICL_FY, S.ICL_CX, S.ICL_CY)
    from difusion_b200.utils.motion_util import Isometry, Rotation
    for f in (0, 40):
        R, t = S.orbit_pose(f, 200)
        rgb, depth = S.render_rgbd(sc, R, t, step=1)
        depth = depth.copy()
        depth[100:140, 200:260] = np.nan                             # a hole, plus the scene's own invalid pixels
        gt = Isometry(q=Rotation(matrix=R), t=t)
        out = []
        for compacting in (False, True):
            trk = SDFTracker(None, args)
            trk.compacting_front_end = compacting
            trk.track_camera(_t(rgb, dev), _t(depth, dev), calib, set_pose=gt)
            out.append((trk.last_processed_pc, trk.last_colored_pcd))
        (pa, ca), (pb, cb) = out
        assert pa[0].shape == pb[0].shape and pa[0].size(0) > 15000
        assert torch.equal(pa[0], pb[0]) and torch.equal(pa[1], pb[1])
        assert torch.equal(ca[0], cb[0]) and torch.equal(ca[1], cb[1])


def test_two_ring_normals_equal_one_ring_normals(dev, monkeypatch):
    """estimate_normals searches the inner 3x3x3 block of a radius/2 grid first and only falls back to the outer shell when the 16
    nearest found so far do not prove completeness; the one-ring search on cells of `radius` (DIF_NORMALS_REACH=1) is the yardstick:
    bit-identical normals, on dense frames, on a frame thinned to a tenth (fallback everywhere) and on an un-compacted cloud."""
    from difusion_b200 import synthetic as S
    from difusion_b200.system import ext
    sc = S.scene_S1(0.05)
    for f, thin in ((0, 1), (30, 1), (60, 1), (10, 10), (10, 3)):
        R, t = S.orbit_pose(f, 200)
        pc, _ = S.frame_points(sc, R, t, box=0.0)
        pc = pc[::thin]
        if f == 30:                                                 # NaN rows in place (the masked front end's input)
            pc = pc.copy(); pc[::7] = np.nan
        cloud = _t(pc, dev)
        monkeypatch.setenv("DIF_NORMALS_REACH", "1")
        one = ext.estimate_normals(cloud, 16, 0.1, [0.0, 0.0, 0.0])
        monkeypatch.delenv("DIF_NORMALS_REACH", raising=False)
        two = ext.estimate_normals(cloud, 16, 0.1, [0.0, 0.0, 0.0])
        assert torch.equal(one.view(torch.int32), two.view(torch.int32)), (f, thin)
        assert int((~torch.isnan(two[:, 0])).sum()) > 0.5 * (pc.shape[0] if f != 30 else pc.shape[0] * 6 // 7) or thin == 10
