"""ICP tensor-core kernel: launch time vs number of 128-sample tiles per CTA (fixed cost vs per-tile cost), cold and warm L2."""
import sys
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
for f in range(3):
    R, t = S.orbit_pose(f); pc, nc = S.frame_points(sc, R, t); xw, nw = S.to_world(pc, nc, R, t)
    m.integrate_keyframe(torch.from_numpy(xw).to(dev), torch.from_numpy(nw).to(dev))
obs0 = torch.from_numpy(pc).to(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for tiles_per_cta in (0.5, 1, 2, 3, 4, 8, 16):
    n = int(148 * 128 * tiles_per_cta)
    obs = obs0.repeat((n + obs0.size(0) - 1) // obs0.size(0), 1)[:n].contiguous()
    for grad in (True, False):
        for cold in (True, False):
            for _ in range(3): m.icp_linearize(obs, R, t, np.eye(3), np.zeros(3), 5.0, grad)
            ts = []
            for _ in range(12):
                if cold: flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); m.icp_linearize(obs, R, t, np.eye(3), np.zeros(3), 5.0, grad); e1.record(); torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            print(f"tiles/CTA={tiles_per_cta:4}  n={n:7d} grad={int(grad)} cold={int(cold)}  median {np.median(ts)*1e3:6.1f} us  min {min(ts)*1e3:6.1f} us")
