"""Time dif_icp_linearize (tensor-core vs fp32 SIMT) on an S1 frame against an S1 map."""
import os, sys
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parents[1]; sys.path.insert(0, str(ROOT))
from difusion_b200 import synthetic as S
from difusion_b200.network import utility as net_util
from difusion_b200.system.map import DenseIndexedMap
dev = torch.device("cuda:0")
model, _ = net_util.load_model(str(ROOT / "tests" / "golden" / "weights.npz"))
sc = S.scene_S1(0.05)
m = DenseIndexedMap(model, sc.map_args(), 29, dev)
frames = []
for f in range(3):
    R, t = S.orbit_pose(f); pc, nc = S.frame_points(sc, R, t); xw, nw = S.to_world(pc, nc, R, t)
    m.integrate_keyframe(torch.from_numpy(xw).to(dev), torch.from_numpy(nw).to(dev)); frames.append((pc, R, t))
pc, R, t = frames[-1]; obs = torch.from_numpy(pc).to(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for path in ("tc", "simt"):
    if path == "simt": os.environ["DIF_ICP_PATH"] = "simt"
    else: os.environ.pop("DIF_ICP_PATH", None)
    for grad in (True, False):
        for _ in range(3): m.icp_linearize(obs, R, t, np.eye(3), np.zeros(3), 5.0, grad)
        ts = []
        for _ in range(10):
            flush.zero_(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); o = m.icp_linearize(obs, R, t, np.eye(3), np.zeros(3), 5.0, grad); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        print(f"{path:5s} want_grad={grad}: n={obs.size(0)} valid={int(o[43])}  median {np.median(ts)*1e3:.1f} us  min {min(ts)*1e3:.1f} us  E={float(o[42]):.6f}")
