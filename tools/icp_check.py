"""Per-element error of the tensor-core ICP linearisation (H, g, energy) against the exact-fp32 SIMT kernel and against the
reference fixture, for both tensor-core pipelines (DIF_ICP_V=1: first pipeline), plus timing on an S1 frame (L2 flushed)."""
import os
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from difusion_b200 import synthetic as S                          # noqa: E402
from difusion_b200.network import utility as net_util            # noqa: E402
from difusion_b200.system.map import DenseIndexedMap             # noqa: E402

dev = torch.device("cuda:0")
model, _ = net_util.load_model(str(ROOT / "tests" / "golden" / "weights.npz"))
sc = S.scene_S1(0.05)
m = DenseIndexedMap(model, sc.map_args(), 29, dev)
for f in range(3):
    R, t = S.orbit_pose(f); pc, nc = S.frame_points(sc, R, t); xw, nw = S.to_world(pc, nc, R, t)
    m.integrate_keyframe(torch.from_numpy(xw).to(dev), torch.from_numpy(nw).to(dev))
obs = torch.from_numpy(pc).to(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
xi = np.asarray([0.004, -0.003, 0.002, 0.003, -0.002, 0.001])
from difusion_b200.utils.motion_util import Isometry             # noqa: E402
delta = Isometry.from_twist(xi)
Rd, td = delta.q.rotation_matrix, delta.t


def run(o_, grad=True):
    return m.icp_linearize(o_, R, t, Rd, td, 5.0, grad).cpu().numpy().copy()


def setenv(path, v):
    for k, val in (("DIF_ICP_PATH", path), ("DIF_ICP_V", v)):
        if val is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = val


def rel(a, b):
    return float(np.max(np.abs(a - b) / (1e-4 + 1e-4 * np.abs(b))))


for reps in (1, 16):
    o_ = obs.repeat(reps, 1).contiguous()
    setenv("simt", None); ref = run(o_); ref_ng = run(o_, False)
    for v in (None, "1"):
        setenv(None, v)
        out = run(o_); out_ng = run(o_, False)
        again = run(o_)
        print(f"n={o_.size(0)} pipeline={'v1' if v else 'v2'}: M {out[43]:.0f} vs {ref[43]:.0f} | H err/tol {rel(out[:36], ref[:36]):.3f} (max-norm rel {np.abs(out[:36]-ref[:36]).max()/np.abs(ref[:36]).max():.2e})"
              f" | g err/tol {rel(out[36:42], ref[36:42]):.3f} (max-norm rel {np.abs(out[36:42]-ref[36:42]).max()/np.abs(ref[36:42]).max():.2e})"
              f" | E {out[42]:.8f} vs {ref[42]:.8f} | E(no grad) {out_ng[42]:.8f} vs {ref_ng[42]:.8f} | reproducible {np.array_equal(out, again)}")
    setenv(None, None)

for v in (None, "1"):
    setenv(None, v)
    for grad in (True, False):
        for _ in range(3):
            run(obs, grad)
        ts = []
        for _ in range(20):
            flush.zero_(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); m.icp_linearize(obs, R, t, Rd, td, 5.0, grad); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        print(f"pipeline={'v1' if v else 'v2'} want_grad={grad}: n={obs.size(0)}  median {np.median(ts)*1e3:.1f} us  min {min(ts)*1e3:.1f} us")
setenv(None, None)

# Does the write-flush itself slow the kernel (its dirty lines are written back while the kernel's misses come in)?  Same timing with
# the flush followed by a 256 MiB READ pass (L2 ends up cold AND clean).
flush2 = torch.empty(64 << 20, dtype=torch.float32, device=dev)
for mode in ("write", "write+read"):
    ts = []
    for _ in range(20):
        flush.zero_()
        if mode == "write+read":
            flush2.sum()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); m.icp_linearize(obs, R, t, Rd, td, 5.0, True); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    print(f"pipeline=v2 flush={mode}: median {np.median(ts)*1e3:.1f} us  min {min(ts)*1e3:.1f} us")
big = obs.repeat(16, 1).contiguous()
ts = []
for _ in range(10):
    flush.zero_(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); m.icp_linearize(big, R, t, Rd, td, 5.0, True); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
print(f"pipeline=v2 16 x frame: n={big.size(0)}  median {np.median(ts)*1e3:.1f} us")
