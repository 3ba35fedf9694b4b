"""Event timeline of CTA 0 of the tensor-core ICP kernel (library built with -DDIF_TC_TRACE:
tools/build_variant.sh trace -DDIF_TC_TRACE ; DIF_LIB_PATH=tools/_build/libdifusion_b200_trace.so python tools/icp_trace.py).
Prints (cycle since the kernel's first instruction, warp, event) for slot 0's epilogue warps 0 / 4, slot 1's warp 8, the issuer and a producer."""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from difusion_b200 import _lib, synthetic as S                     # noqa: E402
from difusion_b200.network import utility as net_util            # noqa: E402
from difusion_b200.system.map import DenseIndexedMap             # noqa: E402

dev = torch.device("cuda:0")
L = _lib.lib()
model, _ = net_util.load_model(str(ROOT / "tests" / "golden" / "weights.npz"))
sc = S.scene_S1(0.05)
m = DenseIndexedMap(model, sc.map_args(), 29, dev)
for f in range(3):
    R, t = S.orbit_pose(f); pc, nc = S.frame_points(sc, R, t); xw, nw = S.to_world(pc, nc, R, t)
    m.integrate_keyframe(torch.from_numpy(xw).to(dev), torch.from_numpy(nw).to(dev))
obs = torch.from_numpy(pc).to(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for _ in range(3):
    m.icp_linearize(obs, R, t, np.eye(3), np.zeros(3), 5.0, True)
buf = torch.zeros(148 * 20 * 8, dtype=torch.int64, device=dev)
flush.zero_()
L.dif_debug_tc_timing(buf.data_ptr())
o = m.icp_linearize(obs, R, t, np.eye(3), np.zeros(3), 5.0, True); torch.cuda.synchronize()
L.dif_debug_tc_timing(None)
raw = buf.cpu().numpy().view(np.uint64)
tr = raw[:20 * 1024].reshape(20, 1024)
gt = raw[20 * 1024:20 * 1024 + 2 * 148 + 2].astype(np.int64)
starts, ends, fin, last = gt[0:296:2], gt[1:296:2], gt[296], gt[297]
s0 = starts[starts > 0].min()
print(f"globaltimer (ns since the first CTA started): CTA starts {starts[starts > 0].min() - s0}..{starts.max() - s0}, CTA role ends (before the last-CTA reduction) "
      f"{ends[ends > 0].min() - s0}..{ends.max() - s0} (median {int(np.median(ends[ends > 0])) - s0}), final result written {fin - s0} by CTA {last}")
ph = raw[20 * 1024 + 300:20 * 1024 + 300 + 8 * 148].astype(np.int64).reshape(148, 8)
labels = ["first tile gathered", "F0 complete", "F1 complete", "F3 complete", "B3 complete", "B0 complete", "tile done"]
for k, lab in enumerate(labels):
    col = ph[:, k][ph[:, k] > 0] - s0
    if col.size:
        print(f"  per-CTA {lab:20s}: min {col.min():6d}  median {int(np.median(col)):6d}  p90 {int(np.percentile(col, 90)):6d}  max {col.max():6d} ns")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
tiny = torch.zeros(32, device=dev)
ts = []
for _ in range(20):
    e0.record(); tiny.add_(1.0); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) * 1e3)
print(f"event pair around a tiny elementwise kernel: median {np.median(ts):.1f} us")
names = {200: "kernel-start", 201: "frame-resolved", 202: "all-roles-done", 203: "partials-written", 5: "I.W1 arrived", 6: "I.W2 arrived", 7: "I.W3 arrived", 1: "I.x-ready A", 2: "I.F0-issued A", 3: "I.x-ready B", 4: "I.F0-issued B",
         60: "E.first-tile-gathered", 80: "E.acc F0", 81: "E.acc F1", 82: "E.acc F2", 83: "E.acc F3", 84: "E.heads+exchange done", 85: "E.acc B3", 86: "E.acc B2",
         87: "E.acc B1", 88: "E.acc B0", 90: "E.g0 F0", 91: "E.g0 F1", 92: "E.g0 F2", 93: "E.g3 handed", 95: "E.done B3", 96: "E.done B2", 97: "E.done B1",
         100: "E.g1 F0", 101: "E.g1 F1", 102: "E.g1 F2", 110: "E.tile-done"}
for st in range(1, 8):
    names[10 + 2 * (st - 1)] = f"I.stage{st}-issued A"
    names[11 + 2 * (st - 1)] = f"I.stage{st}-issued B"
ev = []
for w in (0, 4, 8, 16, 17):
    for e in tr[w]:
        if e:
            ev.append((int(e & np.uint64(0xFFFFFFFFFFFF)), w, int(e >> np.uint64(48))))
ev.sort()
print(f"n={obs.size(0)} valid={int(o[43])} events={len(ev)}")
if ev:
    t0 = ev[0][0]
    for c, w, e in ev:
        print(f"{c - t0:7d}  warp {w:2d}  {names.get(e, e)}")
