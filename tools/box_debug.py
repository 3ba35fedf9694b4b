import sys
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parent.parent; sys.path.insert(0, str(ROOT))
from difusion_b200 import synthetic as S
from difusion_b200.system import ext
dev = torch.device("cuda:0")
sc = S.scene_S1(0.05)
R, t = S.orbit_pose(3)
pc, nc = S.frame_points(sc, R, t, box=0.0)
pc, nc = pc[:20000], nc[:20000]
far = np.concatenate([pc, np.asarray([[400.0, 1.0, 2.0]], np.float32)], 0); nfar = np.concatenate([nc, nc[:1]], 0)
T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
for name, (p, n) in (("near", (pc, nc)), ("far", (far, nfar)), ("near again", (pc, nc)), ("far again", (far, nfar))):
    e_p, e_n = S.box_filter(p, n, 0.02)
    o_p, o_n = ext.point_box_filter(T(p), T(n), 0.02)
    same = o_p.shape[0] == e_p.shape[0] and np.array_equal(o_p.cpu().numpy(), e_p)
    print(name, "gpu rows", o_p.shape[0], "numpy rows", e_p.shape[0], "equal", same, "scratch cells", list(ext._box_scratch.values())[0][2])
