"""Where the host time of one frame goes: perf_counter around the segments of the public-API loop (device-resident inputs)."""
import argparse, sys, time
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parent.parent; sys.path.insert(0, str(ROOT))
import bench
from difusion_b200.network import utility as net_util
from difusion_b200.system.map import DenseIndexedMap
from difusion_b200.system.tracker import SDFTracker
from difusion_b200.utils.motion_util import Isometry, Rotation
K = 200
dev = torch.device("cuda:0")
model, _ = net_util.load_model(str(ROOT / "tests" / "golden" / "weights.npz"))
sc, frames = bench.make_frames(K)
poses = [Isometry(q=Rotation(matrix=fr["R"]), t=fr["t"]) for fr in frames]
ident = Isometry()
trk_args = argparse.Namespace(sdf=dict(robust_kernel="huber", robust_k=5.0, subsample=0.5), rgb=None, iter_config=[{"n": 1, "type": [["sdf"]]}])
d = [dict(pc=torch.from_numpy(fr["pc"]).to(dev), xw=torch.from_numpy(fr["xw"]).to(dev), nw=torch.from_numpy(fr["nw"]).to(dev)) for fr in frames]
for rep in range(3):
    m = DenseIndexedMap(model, sc.map_args(), 29, dev, initial_capacity=1 << 19)
    trk = SDFTracker(m, trk_args)
    torch.cuda.synchronize()
    seg = np.zeros(4)
    t_all = time.perf_counter()
    for f in range(K):
        t0 = time.perf_counter()
        if f >= 1:
            out = m.icp_linearize(d[f]["pc"], poses[f].q.rotation_matrix, poses[f].t, ident.q.rotation_matrix, ident.t, 5.0, True)
        t1 = time.perf_counter()
        if f >= 1:
            o = out.cpu().numpy()
        t2 = time.perf_counter()
        m.integrate_keyframe(d[f]["xw"], d[f]["nw"])
        t3 = time.perf_counter()
        _ = m.n_occupied
        t4 = time.perf_counter()
        seg += [t1 - t0, t2 - t1, t3 - t2, t4 - t3]
    torch.cuda.synchronize()
    tot = time.perf_counter() - t_all
    print(f"rep {rep}: {1e6 * tot / K:.1f} us/frame | icp enqueue {1e6*seg[0]/K:.1f}  icp readback+sync {1e6*seg[1]/K:.1f}  integrate enqueue {1e6*seg[2]/K:.1f}  n_occupied sync {1e6*seg[3]/K:.1f}")

# ---- inside integrate_keyframe (same statements, timed one by one)
import ctypes
from difusion_b200 import _lib
for rep in range(3):
    m = DenseIndexedMap(model, sc.map_args(), 29, dev, initial_capacity=1 << 19)
    torch.cuda.synchronize()
    seg = np.zeros(6)
    for f in range(K):
        xyz, nrm = d[f]["xw"], d[f]["nw"]
        n = xyz.size(0)
        t0 = time.perf_counter()
        m._retire_stats(block=False)
        t1 = time.perf_counter()
        unq = torch.empty(n, dtype=torch.uint8, device=dev)
        t2 = time.perf_counter()
        if n > m._scratch_points:
            m._scratch_points = max(n, 1 << 15)
            m._scratch = torch.empty(m._L.dif_integrate_scratch_bytes(m._scratch_points), dtype=torch.uint8, device=dev)
        view = m._view(); st = _lib.stream_ptr(dev)
        t3 = time.perf_counter()
        _lib.check(m._L.dif_integrate(ctypes.byref(view), m._prep.encoder.data_ptr(), xyz.data_ptr(), nrm.data_ptr(), n, None, _lib.ptr(unq), m._persist.data_ptr(),
                                      m._persist.numel(), m._scratch.data_ptr(), m._scratch.numel(), m._stats_dev.data_ptr(), st), "dif_integrate")
        t4 = time.perf_counter()
        buf, ev = m._stats_ring[m._stats_next]
        buf.copy_(m._stats_dev, non_blocking=True)
        ev.record(torch.cuda.current_stream(dev))
        m._stats_inflight.append((m._stats_next, 7 * n)); m._stats_next = (m._stats_next + 1) % len(m._stats_ring)
        t5 = time.perf_counter()
        _ = m.n_occupied
        t6 = time.perf_counter()
        seg += [t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4, t6 - t5]
    print(f"integrate rep {rep}: retire {1e6*seg[0]/K:.1f}  empty {1e6*seg[1]/K:.1f}  view/stream {1e6*seg[2]/K:.1f}  dif_integrate {1e6*seg[3]/K:.1f}  copy+record {1e6*seg[4]/K:.1f}  sync {1e6*seg[5]/K:.1f}")
