"""One early-frame integrate_keyframe (every PLIVox below encoder_count_th: ~230 k encoder samples) repeated on fresh maps - the launch
`ncu -k regex:encode_tc2_kernel -s 2 -c 1` captures for profiles/r2_ncu_encode_tc_*."""
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent; sys.path.insert(0, str(ROOT))
from difusion_b200 import synthetic as S
from difusion_b200.network import utility as net_util
from difusion_b200.system.map import DenseIndexedMap
dev = torch.device("cuda:0")
model, _ = net_util.load_model(str(ROOT / "tests" / "golden" / "weights.npz"))
sc = S.scene_S1(0.05)
frames = []
for f in range(3):
    R, t = S.orbit_pose(f, 200)
    pc, nc = S.frame_points(sc, R, t)
    xw, nw = S.to_world(pc, nc, R, t)
    frames.append((torch.from_numpy(xw).to(dev), torch.from_numpy(nw).to(dev)))
for rep in range(4):
    m = DenseIndexedMap(model, sc.map_args(), 29, dev, initial_capacity=1 << 18)
    for xw, nw in frames:
        m.integrate_keyframe(xw, nw)
    torch.cuda.synchronize()
    print(rep, m.n_occupied, m.last_integrate_stats)
