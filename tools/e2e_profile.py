"""Host-side profile of the end-to-end frame loop (tracker linearise + integrate through the public API, pinned host inputs).
Prints wall time per frame and the cProfile top functions.  Usage (GPU box):  python tools/e2e_profile.py [frames]
"""
import argparse
import cProfile
import pstats
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from difusion_b200.network import utility as net_util  # noqa: E402
from difusion_b200.system.map import DenseIndexedMap  # noqa: E402
from difusion_b200.system.tracker import SDFTracker  # noqa: E402
from difusion_b200.utils.motion_util import Isometry, Rotation  # noqa: E402


def main():
    K = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    dev = torch.device("cuda:0")
    model, _ = net_util.load_model(str(ROOT / "tests" / "golden" / "weights.npz"))
    sc, frames = bench.make_frames(K)
    poses = [Isometry(q=Rotation(matrix=fr["R"]), t=fr["t"]) for fr in frames]
    ident = Isometry()
    trk_args = argparse.Namespace(sdf=dict(robust_kernel="huber", robust_k=5.0, subsample=0.5), rgb=None, iter_config=[{"n": 1, "type": [["sdf"]]}])
    d = [dict(pc=torch.from_numpy(fr["pc"]).to(dev), xw=torch.from_numpy(fr["xw"]).to(dev), nw=torch.from_numpy(fr["nw"]).to(dev)) for fr in frames]

    def loop(m, trk, sync_every=True):
        for f in range(K):
            if f >= 1:
                trk.compute_sdf_Hg(0, poses[f], ident, d[f]["pc"], no_grad=False)
            m.integrate_keyframe(d[f]["xw"], d[f]["nw"])
            if sync_every:
                _ = m.n_occupied
        torch.cuda.synchronize()

    for rep in range(2):
        m = DenseIndexedMap(model, sc.map_args(), 29, dev, initial_capacity=1 << 19)
        trk = SDFTracker(m, trk_args)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        loop(m, trk)
        dt = time.perf_counter() - t0
        print(f"rep {rep}: {1e6 * dt / K:.1f} us/frame  ({K / dt:.0f} frames/s), device inputs, 2 host syncs per frame")
    m = DenseIndexedMap(model, sc.map_args(), 29, dev, initial_capacity=1 << 19)
    trk = SDFTracker(m, trk_args)
    pr = cProfile.Profile()
    pr.enable()
    loop(m, trk)
    pr.disable()
    pstats.Stats(pr).sort_stats("tottime").print_stats(22)


if __name__ == "__main__":
    main()
