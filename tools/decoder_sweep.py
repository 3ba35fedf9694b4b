"""Decoder batch sweep (BASELINE config 3): samples/s of dif_decode forward for n = 2^14 .. 2^22, latent table from an S1 map."""
import json
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
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
pows = [int(a) for a in sys.argv[1:]] or [14, 16, 18, 20, 22]
out = {}
for p2 in pows:
    n = 1 << p2
    rows = torch.randint(0, table.size(0), (n,), generator=g, dtype=torch.int32).to(dev)
    xyz = (torch.rand(n, 3, generator=g) * 2 - 1).to(dev)
    sdf = torch.empty(n, device=dev); std = torch.empty(n, device=dev)

    def run():
        _lib.check(L.dif_decode(prep.decoder.data_ptr(), table.data_ptr(), 32, rows.data_ptr(), xyz.data_ptr(), n, None, 1.0,
                                sdf.data_ptr(), std.data_ptr(), None, None, _lib.stream_ptr(dev)), "dif_decode")
    for _ in range(3):
        run()
    best = 1e30
    for _ in range(7):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    out[f"2^{p2}"] = dict(ms=best, gsamples_per_s=n / best / 1e6, tflops_algorithmic=n * 98816 / best / 1e9, tflops_issued_3pass=3 * n * 98816 / best / 1e9)
print(json.dumps(out, indent=1))
