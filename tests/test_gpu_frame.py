"""GPU: the one-call / CUDA-graph frame path (dif_frame through FramePipeline) against the separate reference-shaped calls
(SDFTracker.compute_sdf_Hg + DenseIndexedMap.integrate_keyframe) and against the CPU oracle, on the same seeded frames.

The frame path reads the point count and the poses from a device block, strides through interleaved point rows and is replayed
from a captured graph - none of which may change a single bit of the integer map state or of the linearisation.
"""
import argparse

import numpy as np
import pytest
import torch

from conftest import GOLDEN, close

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def model():
    from difusion_b200.network import utility as net_util
    return net_util.load_model(str(GOLDEN / "weights.npz"))[0]


def _frames(n):
    from difusion_b200 import synthetic as S
    sc = S.scene_S1(0.05)
    out = []
    for f in range(n):
        R, t = S.orbit_pose(f, 200)
        pc, nc = S.frame_points(sc, R, t)
        xw, nw = S.to_world(pc, nc, R, t)
        out.append(dict(pc=pc, xw=xw, nw=nw, R=R, t=t))
    return sc, out


def _state(m):
    n = m.n_occupied
    return dict(n=n, indexer=m.indexer.cpu().numpy().copy(), pos=m.latent_vecs_pos.cpu().numpy()[:n].copy(),
                obs=m.voxel_obs_count.cpu().numpy()[:n].copy(), lat=m.latent_vecs.cpu().numpy()[:n].copy())


@pytest.mark.parametrize("use_graph", [False, True])
def test_frame_pipeline_equals_separate_calls(model, dev, use_graph):
    from difusion_b200.system.frame import HDR, ROW, pack_frame
    from difusion_b200.system.map import DenseIndexedMap
    from difusion_b200.system.tracker import SDFTracker
    from difusion_b200.utils.motion_util import Isometry, Rotation
    sc, frames = _frames(5)
    max_n = max(fr["pc"].shape[0] for fr in frames) + 1000            # graph grids are sized by max_points, not by the frame

    # reference-shaped calls, one by one
    ma = DenseIndexedMap(model, sc.map_args(), 29, dev, initial_capacity=1 << 18)
    trk = SDFTracker(ma, argparse.Namespace(sdf=dict(robust_kernel="huber", robust_k=5.0, subsample=0.5), rgb=None,
                                            iter_config=[{"n": 1, "type": [["sdf"]]}]))
    ref = []
    for f, fr in enumerate(frames):
        H = g = E = None
        if f >= 1:
            H, g, E = trk.compute_sdf_Hg(0, Isometry(q=Rotation(matrix=fr["R"]), t=fr["t"]), Isometry(), torch.from_numpy(fr["pc"]).to(dev))
        mask = ma.integrate_keyframe(torch.from_numpy(fr["xw"]).to(dev), torch.from_numpy(fr["nw"]).to(dev))
        ref.append(dict(H=H, g=g, E=E, mask=mask.cpu().numpy(), stats=dict(ma.last_integrate_stats or {}) if ma.n_occupied >= 0 else None))
    sa = _state(ma)

    # one call per frame, device-side frame block, interleaved rows, (optionally) a replayed CUDA graph
    mb = DenseIndexedMap(model, sc.map_args(), 29, dev, initial_capacity=1 << 18)
    pipe = mb.frame_pipeline(max_n, huber_k=5.0, use_graph=use_graph)
    host = torch.zeros(HDR + ROW * max_n, dtype=torch.float32).pin_memory()
    for f, fr in enumerate(frames):
        k = pack_frame(host.numpy(), f + 1, fr["pc"], fr["xw"], fr["nw"], fr["R"], fr["t"])
        buf = pipe.upload(host[:k], fr["pc"].shape[0])
        pipe.launch(buf)
        icp, st = pipe.sync()
        n = fr["pc"].shape[0]
        assert st[8] == f + 1                                          # DIF_STAT_SEQ echoes the frame number
        assert np.array_equal(pipe.unq_mask[:n].cpu().numpy().astype(bool), ref[f]["mask"])
        assert st[4] == ref[f]["stats"]["n_occupied"] and st[2] == ref[f]["stats"]["n_samples"] and st[0] == ref[f]["stats"]["n_kept"]
        if f >= 1:
            # same kernel, same points, same poses (the composite pose is formed with the same fp64 arithmetic on the device); the
            # two maps' latents differ in the last bits (float atomics in the fusion), so the sums agree to ~1e-6, not bit for bit
            H, rH = icp[:36].reshape(6, 6), ref[f]["H"]
            assert np.abs(H - rH).max() <= 1e-4 * np.abs(rH).max()
            assert np.abs(icp[36:42] - ref[f]["g"]).max() <= 1e-4 * np.abs(ref[f]["g"]).max()
            assert close(icp[42], ref[f]["E"], 1e-5) and icp[43] > 0
    if use_graph:
        assert pipe.n_captures == pipe.N_BUF                           # one capture per staging buffer served every frame
    sb = _state(mb)
    assert sa["n"] == sb["n"] and np.array_equal(sa["indexer"], sb["indexer"]) and np.array_equal(sa["pos"], sb["pos"])
    assert np.array_equal(sa["obs"], sb["obs"])
    assert close(sb["lat"], sa["lat"], 1e-6)                           # float atomics: order-dependent in the last bits only

    # reset() in place + replay gives the same map again (what bench.py's repeated passes rely on)
    mb.reset()
    for f, fr in enumerate(frames):
        k = pack_frame(host.numpy(), f + 1, fr["pc"], fr["xw"], fr["nw"], fr["R"], fr["t"])
        pipe.launch(pipe.upload(host[:k], fr["pc"].shape[0]))
        pipe.sync()
    sc2 = _state(mb)
    assert sc2["n"] == sa["n"] and np.array_equal(sc2["indexer"], sa["indexer"]) and np.array_equal(sc2["obs"], sa["obs"])
    assert close(sc2["lat"], sa["lat"], 1e-6)
    if use_graph:
        assert pipe.n_captures == pipe.N_BUF


def test_frame_pipeline_vs_oracle(model, dev):
    """dif_frame end to end against the CPU restatement of the reference (integer state bit-exact, floats at the parity bar)."""
    from difusion_b200.system.frame import HDR, ROW, pack_frame
    from difusion_b200.system.map import DenseIndexedMap
    from oracle import dif_oracle as O
    sc, frames = _frames(3)
    W = O.load_weights_npz(GOLDEN / "weights.npz")
    o = O.OracleMap(W, sc.map_args())
    m = DenseIndexedMap(model, sc.map_args(), 29, dev, initial_capacity=1 << 18)
    max_n = max(fr["pc"].shape[0] for fr in frames)
    pipe = m.frame_pipeline(max_n)
    host = torch.zeros(HDR + ROW * max_n, dtype=torch.float32).pin_memory()
    for f, fr in enumerate(frames):
        k = pack_frame(host.numpy(), f, fr["pc"], fr["xw"], fr["nw"], fr["R"], fr["t"])
        pipe.launch(pipe.upload(host[:k], fr["pc"].shape[0]))
        icp, st = pipe.sync()
        if f >= 1:
            oH, og, oE = O.compute_sdf_Hg(o, fr["R"], fr["t"], np.eye(3), np.zeros(3), fr["pc"], 5.0)
            H = icp[:36].reshape(6, 6)
            assert close(icp[42], oE, 1e-4)
            assert np.abs(H - oH).max() <= 2e-4 * np.abs(oH).max() and np.abs(icp[36:42] - og).max() <= 2e-4 * max(np.abs(og).max(), 1e-3)
        o_mask = o.integrate_keyframe(fr["xw"], fr["nw"])
        assert np.array_equal(pipe.unq_mask[:len(o_mask)].cpu().numpy().astype(bool), o_mask)
        assert st[4] == o.n_occupied
    assert np.array_equal(m.indexer.cpu().numpy(), o.indexer)
    assert np.array_equal(m.voxel_obs_count.cpu().numpy(), o.voxel_obs_count)
    assert close(m.latent_vecs.cpu().numpy(), o.latent_vecs, 1e-4)
