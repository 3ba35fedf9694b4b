"""CPU: the C marching-cubes restatement (oracle/mc_oracle.c) against (1) golden vectors produced by EXECUTING the unmodified
reference CUDA kernel on a B200 (tests/golden/ref_ext_mc_r*.npz, generator tests/golden/make_golden_gpu.py) - bit-exact - and
(2) analytic properties."""
import numpy as np
import pytest

from conftest import GOLDEN

from oracle import mc_oracle


def _sphere_cubes(n_xyz, r, centre, radius, drop=None):
    """One PLIVox per grid cell, cube samples on the (2r)^3 lattice of SURVEY A.10, linear SDF of a sphere (voxel units)."""
    nx, ny, nz = n_xyz
    B = nx * ny * nz
    indexer = np.arange(B, dtype=np.int64).reshape(n_xyz)
    if drop is not None:
        indexer[drop] = -1
    a = -(r // 2) / r
    loc = (np.arange(2 * r) / r + a).astype(np.float32)
    g = np.stack(np.meshgrid(loc, loc, loc, indexing="ij"), -1)               # (2r,2r,2r,3) local coords
    base = np.stack(np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij"), -1).reshape(B, 1, 1, 1, 3)
    p = base + g[None]
    sdf = (np.linalg.norm(p - np.asarray(centre), axis=-1) - radius).astype(np.float32)
    std = np.full_like(sdf, 0.1)
    mapping = np.arange(B, dtype=np.int32)
    return indexer, mapping, sdf, std


def test_sphere_vertices_on_surface_and_closed():
    n_xyz, r = [6, 6, 6], 4
    indexer, mapping, sdf, std = _sphere_cubes(n_xyz, r, (3.0, 3.0, 3.0), 1.7)
    blocks = np.arange(6 ** 3, dtype=np.int64)
    tri, fid, tstd = mc_oracle.marching_cubes_interp(indexer, blocks, mapping, sdf, std, 1 << 20, n_xyz, 10.0)
    assert tri.shape[0] > 500 and tri.shape[1:] == (3, 3)
    rad = np.linalg.norm(tri.reshape(-1, 3) - 3.0, axis=1)
    assert np.abs(rad - 1.7).max() < 0.02                 # linear interpolation error of a curved SDF
    assert np.allclose(tstd, 0.1, atol=1e-6)
    # closed surface: every undirected edge is shared by exactly two triangles
    v = np.round(tri.reshape(-1, 3) * 4096).astype(np.int64)
    key = (v[:, 0] << 40) + (v[:, 1] << 20) + v[:, 2]
    key = key.reshape(-1, 3)
    e = np.concatenate([np.sort(key[:, [0, 1]], 1), np.sort(key[:, [1, 2]], 1), np.sort(key[:, [2, 0]], 1)])
    e = e[e[:, 0] != e[:, 1]]
    _, cnt = np.unique(e, axis=0, return_counts=True)
    assert (cnt == 2).mean() > 0.999
    # flatten ids are the owning PLIVox of each triangle
    cell = np.floor(tri.mean(1)).astype(np.int64)
    own = cell[:, 2] + 6 * cell[:, 1] + 36 * cell[:, 0]
    assert (own == fid).mean() > 0.99


def test_missing_own_block_is_skipped_and_neighbours_renormalise():
    n_xyz, r = [4, 4, 4], 4
    drop = (1, 1, 1)
    indexer, mapping, sdf, std = _sphere_cubes(n_xyz, r, (2.0, 2.0, 2.0), 1.2, drop=drop)
    blocks = np.arange(64, dtype=np.int64)
    tri, fid, _ = mc_oracle.marching_cubes_interp(indexer, blocks, mapping, sdf, std, 1 << 20, n_xyz, 10.0)
    dropped = 1 + 4 * 1 + 16 * 1
    assert not np.any(fid == dropped)
    rad = np.linalg.norm(tri.reshape(-1, 3) - 2.0, axis=1)
    assert np.abs(rad - 1.2).max() < 0.03


def test_max_std_filter_and_overflow():
    n_xyz, r = [4, 4, 4], 4
    indexer, mapping, sdf, std = _sphere_cubes(n_xyz, r, (2.0, 2.0, 2.0), 1.2)
    blocks = np.arange(64, dtype=np.int64)
    full = mc_oracle.marching_cubes_interp(indexer, blocks, mapping, sdf, std, 1 << 20, n_xyz, 10.0)[0].shape[0]
    none = mc_oracle.marching_cubes_interp(indexer, blocks, mapping, sdf, std, 1 << 20, n_xyz, 0.05)[0].shape[0]
    capped = mc_oracle.marching_cubes_interp(indexer, blocks, mapping, sdf, std, 100, n_xyz, 10.0)[0].shape[0]
    assert full > 100 and none == 0 and capped == 100


def test_r5_grid():
    n_xyz, r = [4, 4, 4], 5
    indexer, mapping, sdf, std = _sphere_cubes(n_xyz, r, (2.0, 2.0, 2.0), 1.3)
    tri, _, _ = mc_oracle.marching_cubes_interp(indexer, np.arange(64, dtype=np.int64), mapping, sdf, std, 1 << 20, n_xyz, 10.0)
    rad = np.linalg.norm(tri.reshape(-1, 3) - 2.0, axis=1)
    assert tri.shape[0] > 500 and np.abs(rad - 1.3).max() < 0.02


def _canon(tri, fid, std):
    key = np.concatenate([fid[:, None].astype(np.float64), tri.reshape(len(tri), 9).astype(np.float64)], 1)
    order = np.lexsort(key.T[::-1])
    return tri[order], fid[order], std[order]


@pytest.mark.parametrize("r", [2, 4, 5])
def test_restatement_is_bit_exact_against_the_executed_reference_kernel(r):
    fx = np.load(GOLDEN / f"ref_ext_mc_r{r}.npz")
    for tag in ("all", "flt"):
        tri, fid, std = mc_oracle.marching_cubes_interp(fx["indexer"], fx["blocks"], fx["mapping"], fx["cube_sdf"], fx["cube_std"], 1 << 20,
                                                        fx["n_xyz"].tolist(), float(fx[f"{tag}.max_std"]))
        tri, fid, std = _canon(tri, fid, std)
        assert tri.shape == fx[f"{tag}.tri"].shape and tri.shape[0] > 50
        assert np.array_equal(fid, fx[f"{tag}.fid"])
        assert np.array_equal(tri.view(np.uint32), fx[f"{tag}.tri"].view(np.uint32))
        assert np.array_equal(std.view(np.uint32), fx[f"{tag}.std"].view(np.uint32))


def test_groupby_restatement_against_the_executed_reference_kernel():
    """ext/indexing/indexing.cu:59-109 executed on a B200 vs the oracle's index_add_ restatement (count = L per member row)."""
    import torch
    fx = np.load(GOLDEN / "ref_ext_groupby.npz")
    v, idx, C = torch.from_numpy(fx["values"]), torch.from_numpy(fx["indices"]), int(fx["C"])
    s = torch.zeros(C, v.size(1)).index_add_(0, idx, v).numpy()
    assert np.all(np.abs(s - fx["sum"]) <= 1e-4 + 1e-4 * np.abs(fx["sum"]))
    assert np.array_equal(fx["count"], v.size(1) * np.bincount(fx["indices"], minlength=C))
