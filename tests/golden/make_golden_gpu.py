"""Generates tests/golden/ref_ext_*.npz on a GPU box by EXECUTING THE UNMODIFIED REFERENCE CUDA EXTENSIONS
(oracle/_ref/<name>/<name>.so, built from /root/reference by oracle/build_ref.py) on seeded inputs.

    gpurun -- 'python tests/golden/make_golden_gpu.py gpurun_out/golden'      (then copy the .npz files to tests/golden/)

Fixtures (inputs AND reference outputs are stored, so the CPU tests need neither a GPU nor the reference):
  ref_ext_mc_r{2,4,5}.npz  marching_cubes_sparse_interp (ext/marching_cubes/mc_interp_kernel.cu) on analytic sphere cubes with a missing
                           PLIVox, PLIVoxes left out of the batch, noisy std; outputs for max_std = 10 and 0.15, canonically sorted
  ref_ext_groupby.npz      groupby_sum (ext/indexing/indexing.cu:59-109)
  ref_ext_photo.npz        gradient_xy + rgb_odometry (ext/imgproc/photometric.cu) on two 160x120 synthetic RGB-D views of scene S1
  ref_ext_unproject.npz    unproject_depth (ext/imgproc/imgproc.cu:5-44)
  ref_ext_pcproc.npz       remove_radius_outlier + estimate_normals (ext/pcproc/pcproc.cu) on a noisy 160x120 view of scene S1
The reference ships no vectors of its own, so these executions are the pin for the two CUDA-only ops.
"""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def mc_case(r: int, seed: int = 0):
    """Deterministic MC input: sphere SDF lattice cubes over a small grid (shared with tests/test_oracle_mc.py)."""
    from test_oracle_mc import _sphere_cubes
    n_xyz = [6, 5, 7] if r < 5 else [5, 4, 5]
    c = (np.asarray(n_xyz) / 2.0).tolist()
    indexer, mapping, sdf, std = _sphere_cubes(n_xyz, r, c, 1.9 if r < 5 else 1.6, drop=(2, 2, 3))
    rng = np.random.default_rng(seed)
    std = (std + rng.uniform(0, 0.1, std.shape)).astype(np.float32)
    sdf = (sdf + rng.normal(0, 0.01, sdf.shape)).astype(np.float32)            # neighbouring cubes disagree, like decoder output
    B = sdf.shape[0]
    keep = rng.permutation(B)[: int(B * 0.9)]
    mapping = np.full(B, -1, np.int32)
    mapping[keep] = np.arange(len(keep), dtype=np.int32)
    sdf, std = np.ascontiguousarray(sdf[keep]), np.ascontiguousarray(std[keep])
    blocks = np.sort(rng.choice(B, int(B * 0.8), replace=False)).astype(np.int64)
    blocks = blocks[indexer.reshape(-1)[blocks] != -1]
    return dict(n_xyz=np.asarray(n_xyz), indexer=indexer, blocks=blocks, mapping=mapping, cube_sdf=sdf, cube_std=std)


def photo_case():
    """Two 160x120 views (frames 0 and 6 of the S1 orbit) + the relative pose handed to rgb_odometry, as tracker.py:134-146 builds it."""
    from difusion_b200 import synthetic as S
    sc = S.scene_S1(0.05)
    out = {}
    poses = []
    for tag, f in (("prev", 0), ("cur", 6)):
        R, t = S.orbit_pose(f)
        rgb, depth = S.render_rgbd(sc, R, t, step=4)
        out[f"{tag}_i"] = rgb.mean(-1).astype(np.float32)
        out[f"{tag}_d"] = depth
        poses.append((R, t))
    (R0, t0), (R1, t1) = poses
    # delta = last^-1 . cur (tracker.py:223), then perturbed so that the residuals are not at their minimum
    Rd = R0.T @ R1
    td = R0.T @ (t1 - t0) + np.array([0.004, -0.003, 0.002])
    fx, fy, cx, cy = S.ICL_FX / 4, S.ICL_FY / 4, S.ICL_CX / 4, S.ICL_CY / 4
    K = np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1.0]])
    out.update(intr=np.array([fx, fy, cx, cy], np.float32), K=K, Rd=Rd, td=td,
               krkinv=(K @ Rd @ np.linalg.inv(K)).flatten(), kt=(K @ td).flatten(),
               min_grad_scale=np.float32(1e-5), max_depth_delta=np.float32(0.2))
    return out


def cloud_case(step: int = 4, noise: float = 0.004, seed: int = 11):
    """The point cloud track_camera hands to the kd-tree ops (tracker.py:92-104): unprojected sub-sampled depth, (N,4) rows."""
    from difusion_b200 import synthetic as S
    sc = S.scene_S1(0.05)
    R, t = S.orbit_pose(30)
    depth, _ = S.render_depth(sc, R, t, noise_sigma=noise, seed=seed, step=step)
    rng = np.random.default_rng(seed)
    depth = depth.copy()
    fly = rng.random(depth.shape) < 0.01                      # isolated flying pixels: what the radius filter is for
    depth[fly] = depth[fly] * np.float32(0.8)
    h, w = depth.shape
    fx, fy, cx, cy = (np.float32(S.ICL_FX / step), np.float32(S.ICL_FY / step), np.float32(S.ICL_CX / step), np.float32(S.ICL_CY / step))
    uu, vv = np.meshgrid(np.arange(w, dtype=np.float32), np.arange(h, dtype=np.float32))
    pc = np.stack([(uu - cx) / fx * depth, (vv - cy) / fy * depth, depth, np.zeros_like(depth)], -1).reshape(-1, 4)
    return np.ascontiguousarray(pc[~np.isnan(pc[:, 0])].astype(np.float32))


def canon(tri, fid, std):
    """Canonical order: by PLIVox id, then by the 9 vertex coordinates (exact floats)."""
    key = np.concatenate([fid[:, None].astype(np.float64), tri.reshape(len(tri), 9).astype(np.float64)], 1)
    order = np.lexsort(key.T[::-1])
    return tri[order], fid[order], std[order]


def main():
    out = Path(sys.argv[1] if len(sys.argv) > 1 else ROOT / "gpurun_out" / "golden")
    out.mkdir(parents=True, exist_ok=True)
    from oracle import build_ref
    dev = torch.device("cuda:0")
    mc = build_ref.load_module("marching_cubes")
    ix = build_ref.load_module("indexing")
    for r in (2, 4, 5):
        c = mc_case(r)
        t = {k: torch.from_numpy(v).to(dev) for k, v in c.items() if k != "n_xyz"}
        res = {}
        for tag, max_std in (("all", 10.0), ("flt", 0.15)):
            tri, fid, std = mc.marching_cubes_sparse_interp(t["indexer"], t["blocks"], t["mapping"], t["cube_sdf"], t["cube_std"], 1 << 20,
                                                            c["n_xyz"].tolist(), max_std)
            tri, fid, std = canon(tri.cpu().numpy(), fid.cpu().numpy(), std.cpu().numpy())
            res.update({f"{tag}.tri": tri, f"{tag}.fid": fid, f"{tag}.std": std, f"{tag}.max_std": np.float32(max_std)})
            print(f"mc r={r} max_std={max_std}: {tri.shape[0]} triangles")
        np.savez_compressed(out / f"ref_ext_mc_r{r}.npz", r=r, **c, **res)
    g = torch.Generator().manual_seed(0)
    v = torch.randn(5000, 29, generator=g)
    idx = torch.randint(0, 37, (5000,), generator=g)
    s, cnt = ix.groupby_sum(v.to(dev), idx.to(dev), 40)
    np.savez_compressed(out / "ref_ext_groupby.npz", values=v.numpy(), indices=idx.numpy(), C=40, sum=s.cpu().numpy(), count=cnt.cpu().numpy())
    print("groupby_sum:", s.shape, cnt.dtype, int(cnt.sum()))
    im = build_ref.load_module("imgproc")
    c = photo_case()
    t = {k: torch.from_numpy(np.ascontiguousarray(c[k])).to(dev) for k in ("prev_i", "prev_d", "cur_i", "cur_d")}
    grad = im.gradient_xy(t["cur_i"])
    f_img, J_img = im.rgb_odometry(t["prev_i"], t["prev_d"], t["cur_i"], t["cur_d"], grad, c["intr"].tolist(), c["krkinv"].tolist(), c["kt"].tolist(),
                                   float(c["min_grad_scale"]), float(c["max_depth_delta"]), True)
    f_only, = im.rgb_odometry(t["prev_i"], t["prev_d"], t["cur_i"], t["cur_d"], grad, c["intr"].tolist(), c["krkinv"].tolist(), c["kt"].tolist(),
                              float(c["min_grad_scale"]), float(c["max_depth_delta"]), False)
    f_np, J_np = f_img.cpu().numpy(), J_img.cpu().numpy()
    assert np.array_equal(np.isnan(f_np), np.isnan(f_only.cpu().numpy()))
    J_np[np.isnan(f_np)] = np.nan                          # uninitialised memory in the reference: not part of the contract
    np.savez_compressed(out / "ref_ext_photo.npz", **c, grad=grad.cpu().numpy(), f=f_np, J=J_np)
    print("photo: valid pixels", int((~np.isnan(f_np)).sum()), "of", f_np.size)
    depth = c["cur_d"]
    pc = im.unproject_depth(torch.from_numpy(depth).to(dev), float(c["intr"][0]), float(c["intr"][1]), float(c["intr"][2]), float(c["intr"][3])).cpu().numpy()
    pc[np.isnan(depth)] = np.nan                           # only x = NaN is written for invalid pixels (imgproc.cu:21)
    np.savez_compressed(out / "ref_ext_unproject.npz", depth=depth, intr=c["intr"], pc=pc)
    print("unproject:", pc.shape)
    pp = build_ref.load_module("pcproc")
    cloud = cloud_case()
    d_cloud = torch.from_numpy(cloud).to(dev)
    res = {}
    for tag, radius in (("r5", 0.05), ("r8", 0.08)):           # 0.05 is the tracker's setting; at 160x120 it rejects most points
        mask = pp.remove_radius_outlier(d_cloud, 16, radius)
        kept = d_cloud[mask].contiguous()
        normals = pp.estimate_normals(kept, 16, 2 * radius, [0.0, 0.0, 0.0])
        res[f"{tag}.radius"] = np.float32(radius)
        res[f"{tag}.mask"] = mask.cpu().numpy()
        res[f"{tag}.normals"] = normals.cpu().numpy()
        print(f"pcproc radius {radius}: kept {int(mask.sum())} of {cloud.shape[0]}, NaN normals {int(torch.isnan(normals[:, 0]).sum())}")
    np.savez_compressed(out / "ref_ext_pcproc.npz", cloud=cloud, **res)


if __name__ == "__main__":
    main()
