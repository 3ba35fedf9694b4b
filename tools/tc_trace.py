"""Event timeline of CTA 0 of the tensor-core decoder (library built with -DDIF_TC_TRACE; tools/build_variant.sh trace -DDIF_TC_TRACE).
Prints, per warp of interest, the first events after a steady-state offset as (cycle relative to the window start, event)."""
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
n = 1 << 20
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
run(); torch.cuda.synchronize()
L.dif_debug_tc_timing(None)
t = buf.cpu().numpy().view(np.uint64)[:20 * 1024].reshape(20, 1024)
names = {1: "I.x-ready", 2: "I.L0-issued", 80: "E.acc0", 81: "E.acc1", 82: "E.acc2", 83: "E.acc3", 90: "E.g0.L0", 91: "E.g0.L1", 92: "E.g0.L2",
         100: "E.g1.L0", 101: "E.g1.L1", 102: "E.g1.L2", 110: "E.heads-done", 120: "P.x-free", 121: "P.x-ready", 122: "P.st0", 123: "P.st1", 124: "P.resolved-issued", 125: "P.st2", 126: "P.st3"}
for layer in (1, 2, 3):
    for gg in range(4):
        names[16 + 4 * layer + gg] = f"I.L{layer}g{gg}-ready"
        names[48 + 4 * layer + gg] = f"I.L{layer}g{gg}-issued"
ev = []
for w in (0, 4, 16, 17):                   # slot 0: epilogue quad 0 half 0, half 1; issuer; producer
    for e in t[w]:
        if e:
            ev.append((int(e & np.uint64(0xFFFFFFFFFFFF)), w, int(e >> np.uint64(48))))
ev.sort()
t_issuer = [c for c, w, e in ev if w == 16 and e == 1]
start = t_issuer[6] if len(t_issuer) > 8 else ev[0][0]            # 7th tile of slot 0: steady state
end = t_issuer[8] if len(t_issuer) > 8 else ev[-1][0]
print(f"window: tiles 6..7 of slot 0, {end - start} cycles for 2 tiles")
for c, w, e in ev:
    if start <= c <= end:
        print(f"{c - start:7d}  warp {w:2d}  {names.get(e, e)}")
