"""Phase breakdown of the tensor-core decoder kernel (cycles per warp role), via dif_debug_tc_timing."""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from difusion_b200 import _lib                                     # noqa: E402
from difusion_b200.network import utility as net_util            # noqa: E402

dev = torch.device("cuda:0")
L = _lib.lib()
model, _ = net_util.load_model(str(ROOT / "tests" / "golden" / "weights.npz"))
prep = net_util.prepared_for(model, dev)
g = torch.Generator().manual_seed(0)
table = torch.zeros(23000, 32)
table[:, :29] = torch.randn(23000, 29, generator=g) * 0.2          # 128-byte rows, as the map stores them
table = table.to(dev)
n = 1 << int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 22
rows = torch.randint(0, table.size(0), (n,), generator=g, dtype=torch.int32).to(dev)
xyz = (torch.rand(n, 3, generator=g) * 2 - 1).to(dev)
sdf = torch.empty(n, device=dev); std = torch.empty(n, device=dev)


def run():
    _lib.check(L.dif_decode(prep.decoder.data_ptr(), table.data_ptr(), 32, rows.data_ptr(), xyz.data_ptr(), n, None, 1.0,
                            sdf.data_ptr(), std.data_ptr(), None, None, _lib.stream_ptr(dev)), "dif_decode")


for _ in range(3):
    run()
buf = torch.zeros(148 * 20 * 8, dtype=torch.int64, device=dev)
L.dif_debug_tc_timing(buf.data_ptr())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); run(); e1.record(); torch.cuda.synchronize()
L.dif_debug_tc_timing(None)
t = buf.cpu().numpy().reshape(148, 20, 8).astype(np.float64)
tiles_per_cta = n / 128 / 148
print(f"n={n}  kernel {e0.elapsed_time(e1):.3f} ms (with timing code)  tiles/CTA {tiles_per_cta:.1f}")
mma = t[:, 16, :].mean(0)
print(f"MMA thread : weight load {mma[0]:.0f}  waiting {mma[1]:.0f}  issuing {mma[2]:.0f}   (cycles per CTA);  per tile: wait {mma[1]/tiles_per_cta:.0f} issue {mma[2]/tiles_per_cta:.0f}")
ep = t[:, :16, :].mean((0, 1))
tp = tiles_per_cta / 2
print(f"epilogue warp (avg of 16): loop {ep[7]/tp:.0f}  wait-acc {ep[4]/tp:.0f}  convert(3 layers) {ep[5]/tp:.0f}  last+heads {ep[6]/tp:.0f}  cycles per tile")
print(f"sum per tile per slot: {(ep[4]+ep[5]+ep[6]+ep[7])/tp:.0f}")
pr = t[:, 17:19, :].mean((0, 1))
print(f"producer warp per tile: resolve+prefetch {pr[0]/tp:.0f}  wait-x-free {pr[1]/tp:.0f}  first pair {pr[2]/tp:.0f}  rest {pr[3]/tp:.0f}  fence+arrive {pr[4]/tp:.0f}  | loads {pr[5]/tp:.0f} stores {pr[6]/tp:.0f}")
