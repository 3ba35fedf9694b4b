import sys
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parent.parent; sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from difusion_b200 import synthetic as S
from difusion_b200.network import utility as net_util
from difusion_b200.system.map import DenseIndexedMap
dev = torch.device("cuda:0")
GOLDEN = ROOT / "tests" / "golden"
fx = np.load(GOLDEN / "s0_optimize.npz")
sc = S.scene_S0(); args = sc.map_args()
args.encoder_count_th = float(fx["encoder_count_th"]); args.optim_n_iters, args.code_regularization, args.code_reg_lambda = int(fx["n_iters"]), True, float(fx["code_reg_lambda"])
model, _ = net_util.load_model(str(GOLDEN / "weights.npz"))
m = DenseIndexedMap(model, args, 29, dev)
m.optim_noise_fn = S.ReproducibleNoise()
rec = {}
orig = m.optimize_latent_rows
def spy(lat, inv, sdf, rel):
    rec["inv"] = inv.clone(); rec["lat0"] = lat.clone(); rec["rel"] = rel.clone()
    return orig(lat, inv, sdf, rel)
m.optimize_latent_rows = spy
R, t = S.yaw_pose(float(fx["f0.yaw"])); pc, nc = S.frame_points(sc, R, t); xw, nw = S.to_world(pc, nc, R, t)
m.integrate_keyframe(torch.from_numpy(xw).to(dev), torch.from_numpy(nw).to(dev), do_optimize=True)
n = int(fx["f0.n_occupied"])
lat, ref = m.latent_vecs.cpu().numpy()[:n], fx["f0.latent"]
d = np.abs(lat - ref).max(1)
bad = np.nonzero(d > 1e-4)[0]
print("rows off:", len(bad), "of", n, "max", d.max())
cnt = np.bincount(rec["inv"].cpu().numpy())
opt_rows = np.nonzero(fx["f0.optimized"])[0]
print("samples per optimised row: min/median/max", cnt.min(), np.median(cnt), cnt.max(), "total", cnt.sum())
pos = m.latent_vecs_pos.cpu().numpy()[:n]
for r in bad[:12]:
    k = np.searchsorted(opt_rows, r)
    print("row", r, "diff", d[r], "obs", fx["f0.obs_count"][r], "samples", cnt[k] if k < len(cnt) else None, "cell", np.unravel_index(pos[r], (32, 32, 32)), "|lat|", np.linalg.norm(ref[r]))
print("typical good row samples:", [int(cnt[np.searchsorted(opt_rows, r)]) for r in opt_rows[:12] if d[r] <= 1e-4])
