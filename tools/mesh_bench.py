"""BASELINE config 4: full-scene mesh extraction at 1 cm (5 cm PLIVoxes, voxel_resolution=5) on a ~50k-PLIVox map.
Scene S2 of SURVEY 8(d): Fibonacci-lattice sphere R=3.15 m, 3 M points + inward normals, integrated in 10 calls."""
import json, sys, time
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parents[1]; sys.path.insert(0, str(ROOT))
from difusion_b200 import _lib, synthetic as S
from difusion_b200.network import utility as net_util
from difusion_b200.system.map import DenseIndexedMap
from difusion_b200.system import ext

dev = torch.device("cuda:0")
model, _ = net_util.load_model(str(ROOT / "tests" / "golden" / "weights.npz"))
R = float(sys.argv[1]) if len(sys.argv) > 1 else 3.15
n_pts = int(float(sys.argv[2])) if len(sys.argv) > 2 else 3_000_000
sc = S.scene_S2(R)
pts, nrm = S.s2_sphere_points(R, n_pts)
m = DenseIndexedMap(model, sc.map_args(), 29, dev, initial_capacity=1 << 18)
t0 = time.perf_counter()
for c in range(10):
    sl = slice(c * n_pts // 10, (c + 1) * n_pts // 10)
    m.integrate_keyframe(torch.from_numpy(pts[sl]).to(dev), torch.from_numpy(nrm[sl]).to(dev))
nocc = m.n_occupied
torch.cuda.synchronize()
print(f"integrated {n_pts} points in 10 calls: {time.perf_counter() - t0:.3f} s wall; n_occupied={nocc}, observed={(m.voxel_obs_count[:nocc] > 4).sum().item()}")

def ev(): return torch.cuda.Event(enable_timing=True)
res = {}
for rep in range(3):
    e = [ev() for _ in range(4)]
    e[0].record()
    focused, mapping, cs, cd, slots, cnt = m.mesh_cubes(5, fast=True, updated_vec_id=None)
    e[1].record()
    tri, fid, tstd = ext.marching_cubes_interp(m.indexer.view(m.n_xyz), focused, mapping, cs, cd, int(12e6), m.n_xyz, 0.15)
    e[2].record()
    torch.cuda.synchronize()
    n_low, n_high = cnt.tolist()
    res = dict(B=int(cs.size(0)), K=int(focused.numel()), n_low=n_low, n_high=n_high, triangles=int(tri.size(0)),
               select_decode_ms=e[0].elapsed_time(e[1]), marching_cubes_ms=e[1].elapsed_time(e[2]))
    print(json.dumps(res))
# MC alone with events around the kernel only
L = _lib.lib()
k0, k1 = ev(), ev()
for _ in range(2):
    L.dif_profile_hook(3, k0.cuda_event if hasattr(k0, "cuda_event") else None, None)
k0.record(); k1.record()
L.dif_profile_hook(3, k0.cuda_event, k1.cuda_event)
tri, fid, tstd = ext.marching_cubes_interp(m.indexer.view(m.n_xyz), focused, mapping, cs, cd, int(12e6), m.n_xyz, 0.15)
torch.cuda.synchronize()
mc_ms = k0.elapsed_time(k1)
T = tri.size(0); B = cs.size(0)
alg_bytes = 8 * 1000 * B + 8 * focused.numel() + 56 * T
print(f"marching_cubes_kernel: {mc_ms*1e3:.1f} us for K={focused.numel()} PLIVoxes, T={T} triangles; algorithmic {alg_bytes/1e6:.1f} MB -> {alg_bytes/mc_ms/1e6:.1f} GB/s")
# the reference's own kernel (oracle/_ref, unmodified) on the same cubes, timed the same way
from oracle import build_ref
if build_ref.available("marching_cubes"):
    rmc = build_ref.load_module("marching_cubes")
    args = (m.indexer.view(m.n_xyz), focused, mapping, cs, cd, int(12e6), m.n_xyz, 0.15)
    for _ in range(2):
        rmc.marching_cubes_sparse_interp(*args)
    ts = []
    for _ in range(5):
        e0, e1 = ev(), ev(); e0.record(); rt = rmc.marching_cubes_sparse_interp(*args); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    print(f"reference marching_cubes_sparse_interp (allocation + kernel + sync + trim, as shipped): {min(ts)*1e3:.1f} us, T={rt[0].size(0)}")
    ts = []
    for _ in range(5):
        e0, e1 = ev(), ev(); e0.record(); ot = ext.marching_cubes_interp(*args); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    print(f"ours     system.ext.marching_cubes_interp (same call shape):                           {min(ts)*1e3:.1f} us, T={ot[0].size(0)}")
# properties: vertices on the sphere, std filter respected
v = tri.reshape(-1, 3) * sc.voxel_size + torch.tensor(sc.bound_min, device=dev)
rad = v.norm(dim=1)
print(f"vertex radius: mean {rad.mean().item():.4f} m, max |r-R| {(rad - R).abs().max().item():.4f} m (voxel 0.05 m); max vertex std {tstd.max().item():.3f} (<= 0.15)")
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    full = m.extract_mesh(5, int(12e6), max_std=0.15, no_cache=True)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    print(f"extract_mesh no_cache (select + decode + MC + device cache merge, mesh left on the device): {full.n_triangles} triangles, {1e3 * (t1 - t0):.2f} ms wall")
t0 = time.perf_counter(); v = full.vertices; t1 = time.perf_counter()
print(f"download of the mesh on demand: {v.shape[0]} vertices, {1e3 * (t1 - t0):.1f} ms")
# incremental: touch a patch, re-extract (device-side merge of the cache); three rounds (the first one sizes the workspaces)
for rep in range(3):
    sl = slice(rep * 30000, (rep + 1) * 30000)
    m.integrate_keyframe(torch.from_numpy(pts[sl]).to(dev), torch.from_numpy(nrm[sl]).to(dev))
    torch.cuda.synchronize(); t0 = time.perf_counter()
    inc = m.extract_mesh(5, int(12e6), max_std=0.15)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    print(f"incremental extract_mesh after one 30k-point frame: {inc.n_triangles} triangles in cache, {1e3 * (t1 - t0):.2f} ms wall")
