"""GPU: the product kernels against the UNMODIFIED reference CUDA extensions executed on the same device.

oracle/build_ref.py compiles the reference's own sources (ext/marching_cubes, ext/indexing) into oracle/_ref/ in the build
container; the .so files travel to the GPU box with the snapshot.  These tests load them next to libdifusion_b200.so and feed
both the same CUDA tensors.  Bars: marching cubes - identical triangle multiset, vertices / std / ids BIT-EXACT; groupby_sum -
counts exact, sums within 1e-4 (both sides use float atomics in arbitrary order).
"""
import numpy as np
import pytest
import torch

from conftest import close

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return torch.device("cuda:0")


def _ref(name):
    from oracle import build_ref
    if not build_ref.available(name):
        pytest.skip(f"oracle/_ref/{name} not built (python oracle/build_ref.py needs /root/reference)")
    return build_ref.load_module(name)


def _canon(tri, fid, std):
    tri, fid, std = tri.cpu().numpy(), fid.cpu().numpy(), std.cpu().numpy()
    key = np.concatenate([fid[:, None].astype(np.float64), tri.reshape(len(tri), 9).astype(np.float64)], 1)
    order = np.lexsort(key.T[::-1])
    return tri[order], fid[order], std[order]


def _assert_same_mesh(ours, ref):
    a, b = _canon(*ours), _canon(*ref)
    assert a[0].shape == b[0].shape and a[0].shape[0] > 0
    assert np.array_equal(a[1], b[1])
    assert np.array_equal(a[0].view(np.uint32), b[0].view(np.uint32)), f"vertices differ, max |d| = {np.abs(a[0] - b[0]).max():.3e}"
    assert np.array_equal(a[2].view(np.uint32), b[2].view(np.uint32)), f"vertex std differs, max |d| = {np.abs(a[2] - b[2]).max():.3e}"


@pytest.mark.parametrize("r", [2, 4, 5])
def test_marching_cubes_vs_reference_kernel_synthetic(dev, r):
    from difusion_b200.system import ext
    from golden.make_golden_gpu import mc_case
    mc = _ref("marching_cubes")
    c = mc_case(r, seed=r)
    t = {k: torch.from_numpy(v).to(dev) for k, v in c.items() if k != "n_xyz"}
    n_xyz = c["n_xyz"].tolist()
    for max_std in (10.0, 0.15):
        args = (t["indexer"], t["blocks"], t["mapping"], t["cube_sdf"], t["cube_std"], 1 << 20, n_xyz, max_std)
        _assert_same_mesh(ext.marching_cubes_interp(*args), mc.marching_cubes_sparse_interp(*args))


def test_marching_cubes_vs_reference_kernel_on_a_fused_map(dev):
    """Real decoder cubes (scene S1, three frames, 5 cm PLIVoxes, 1 cm mesh): every corner blends up to 8 disagreeing PLIVoxes."""
    from conftest import GOLDEN
    from difusion_b200 import synthetic as S
    from difusion_b200.network import utility as net_util
    from difusion_b200.system import ext
    from difusion_b200.system.map import DenseIndexedMap
    mc = _ref("marching_cubes")
    model, _ = net_util.load_model(str(GOLDEN / "weights.npz"))
    sc = S.scene_S1(0.05)
    m = DenseIndexedMap(model, sc.map_args(), 29, dev)
    for f in (0, 20, 40):
        R, tt = S.orbit_pose(f)
        pc, nc = S.frame_points(sc, R, tt)
        xw, nw = S.to_world(pc, nc, R, tt)
        m.integrate_keyframe(torch.from_numpy(xw).to(dev), torch.from_numpy(nw).to(dev))
    for r in (4, 5):
        focused, mapping, cs, cd, _, _ = m.mesh_cubes(r, fast=True)
        args = (m.indexer.view(m.n_xyz), focused, mapping, cs, cd, int(6e6), m.n_xyz, 0.15)
        ours, ref = ext.marching_cubes_interp(*args), mc.marching_cubes_sparse_interp(*args)
        assert ours[0].shape[0] > 50_000
        _assert_same_mesh(ours, ref)


def test_marching_cubes_vs_reference_kernel_config4_full_size(dev):
    """BASELINE config 4 at its full size: scene S2 (sphere, 3 M points, ~147 k PLIVoxes at 5 cm), 1 cm mesh (voxel_resolution 5),
    ~3.8 M triangles: identical triangle multiset, vertices / std / ids bit-exact against the executed reference kernel."""
    from conftest import GOLDEN
    from difusion_b200 import synthetic as S
    from difusion_b200.network import utility as net_util
    from difusion_b200.system import ext
    from difusion_b200.system.map import DenseIndexedMap
    mc = _ref("marching_cubes")
    model, _ = net_util.load_model(str(GOLDEN / "weights.npz"))
    sc = S.scene_S2()
    pts, nrm = S.s2_sphere_points()
    m = DenseIndexedMap(model, sc.map_args(), 29, dev, initial_capacity=1 << 18)
    for c in range(10):
        sl = slice(c * len(pts) // 10, (c + 1) * len(pts) // 10)
        m.integrate_keyframe(torch.from_numpy(pts[sl]).to(dev), torch.from_numpy(nrm[sl]).to(dev))
    assert m.n_occupied >= 50_000                                # "full scene, >= 50 k blocks" (BASELINE.json configs[3])
    focused, mapping, cs, cd, _, _ = m.mesh_cubes(5, fast=True)
    args = (m.indexer.view(m.n_xyz), focused, mapping, cs, cd, int(12e6), m.n_xyz, 0.15)
    ours, ref = ext.marching_cubes_interp(*args), mc.marching_cubes_sparse_interp(*args)
    assert ours[0].shape[0] > 3_000_000
    _assert_same_mesh(ours, ref)
    # and the property the scene was chosen for: the mesh sits on the sphere (mean radius to the millimetre; single vertices up to
    # two 5 cm PLIVoxes off where the blend of neighbouring PLIVoxes disagrees - the reference's mesh has the same vertices)
    v = ours[0].reshape(-1, 3) * sc.voxel_size + torch.tensor(sc.bound_min, device=dev)
    rad = v.norm(dim=1)
    assert abs(float(rad.mean()) - 3.15) < 2e-3 and float((rad - 3.15).abs().max()) < 0.15


def test_groupby_sum_vs_reference_kernel(dev):
    from difusion_b200.system import ext
    ix = _ref("indexing")
    g = torch.Generator().manual_seed(1)
    for n, C in ((5000, 40), (200_000, 9000), (7, 3)):
        v = torch.randn(n, 29, generator=g).to(dev)
        idx = torch.randint(0, C - 1, (n,), generator=g).to(dev)
        s, c = ext.groupby_sum(v, idx, C)
        rs, rc = ix.groupby_sum(v, idx, C)
        assert s.shape == rs.shape and c.dtype == rc.dtype and torch.equal(c, rc)
        assert close(s.cpu().numpy(), rs.cpu().numpy(), 1e-4)
