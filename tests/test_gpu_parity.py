"""GPU: the CUDA path (through the reference-shaped Python API -> C ABI -> sm_100a kernels) against the golden fixtures
produced by the unmodified reference, and against the oracle on seeded inputs.

Parity bar (BASELINE.json north_star): integer / index state bit-exact; floats |a-b| <= 1e-4 + 1e-4|b|.
"""
import numpy as np
import pytest
import torch

from conftest import GOLDEN, close, fixture_args, frac_off, hg_errors

pytestmark = pytest.mark.gpu
TOL = 1e-4


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def model():
    from difusion_b200.network import utility as net_util
    return net_util.load_model(str(GOLDEN / "weights.npz"))[0]


def _t(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def test_native_library_is_loaded():
    from difusion_b200 import _lib
    assert _lib.lib().dif_abi_version() == _lib.ABI_VERSION
    assert "libdifusion_b200.so" in open("/proc/self/maps").read()


def test_decoder_kat(golden, model, dev):
    from difusion_b200.network import utility as net_util
    fx = golden["decoder_kat"]
    lat, xyz = _t(fx["latent"], dev), _t(fx["xyz"], dev).requires_grad_(True)
    sdf, std = net_util.forward_model(model.decoder, latent_input=lat, xyz_input=xyz, no_detach=True)
    assert sdf.shape == (4096, 1) and std.shape == (4096, 1)
    assert close(sdf.detach().cpu().numpy()[:, 0], fx["sdf"], TOL) and close(std.detach().cpu().numpy()[:, 0], fx["std"], TOL)
    g = torch.autograd.grad(sdf.sum(), xyz, retain_graph=True)[0].cpu().numpy()
    assert frac_off(g, fx["dsdf_dxyz"]) < 2e-3                  # ReLU-kink flips (see tests/golden/make_golden.py)
    g2 = torch.autograd.grad(std.sum(), xyz)[0].cpu().numpy()
    assert frac_off(g2, fx["dstd_dxyz"]) < 2e-3
    # network_input form + detach semantics (utility.py:78-80,118-119)
    out = net_util.forward_model(model.decoder, network_input=torch.cat([lat, xyz.detach()], 1))
    assert not out[0].requires_grad and close(out[0].cpu().numpy()[:, 0], fx["sdf"], TOL)


def test_encoder_kat(golden, model, dev):
    fx = golden["encoder_kat"]
    out = model.encoder(_t(fx["xyzn"], dev)).cpu().numpy()
    assert close(out, fx["latent"], TOL)


def test_decoder_ragged_and_padding(golden, model, dev):
    """sizes that are not a multiple of the tile, n=1, and negative rows (padding)."""
    from difusion_b200 import _lib
    from difusion_b200.network import utility as net_util
    fx = golden["decoder_kat"]
    prep = net_util.prepared_for(model, dev)
    for n in (1, 31, 33, 1000):
        sdf, std = net_util.forward_model(model.decoder, latent_input=_t(fx["latent"][:n], dev), xyz_input=_t(fx["xyz"][:n], dev))
        assert close(sdf.cpu().numpy()[:, 0], fx["sdf"][:n], TOL) and close(std.cpu().numpy()[:, 0], fx["std"][:n], TOL)
    n = 100
    rows = torch.arange(n, dtype=torch.int32, device=dev)
    rows[::3] = -1
    lat, xyz = _t(fx["latent"][:n], dev), _t(fx["xyz"][:n], dev)
    sdf = torch.full((n,), 7.0, device=dev); std = torch.full((n,), 7.0, device=dev)
    _lib.check(_lib.lib().dif_decode(prep.decoder.data_ptr(), lat.data_ptr(), lat.stride(0), rows.data_ptr(), xyz.data_ptr(), n, None, -1.0,
                                     sdf.data_ptr(), std.data_ptr(), None, None, _lib.stream_ptr(dev)), "dif_decode")
    live = (rows >= 0).cpu().numpy()
    assert close(-sdf.cpu().numpy()[live], fx["sdf"][:n][live], TOL) and np.all(sdf.cpu().numpy()[~live] == 0)


@pytest.mark.parametrize("name", ["s0_map", "s0_freeze", "s1_map"])
def test_map_against_reference_fixture(golden, model, dev, name):
    from difusion_b200.system.map import DenseIndexedMap
    from difusion_b200.system.tracker import SDFTracker
    from difusion_b200.utils.motion_util import Isometry, Rotation
    import argparse
    fx = golden[name]
    m = DenseIndexedMap(model, fixture_args(fx), 29, dev)
    for f in range(int(fx["n_frames"])):
        mask = m.integrate_keyframe(_t(fx[f"f{f}.xyz"], dev), _t(fx[f"f{f}.normal"], dev))
        assert mask.dtype == torch.bool and np.array_equal(np.packbits(mask.cpu().numpy()), fx[f"f{f}.unq_mask"])
        nocc = m.n_occupied
        assert nocc == int(fx[f"f{f}.n_occupied"])
        assert m.latent_vecs.size(0) == int(fx[f"f{f}.capacity"]) == m.latent_vecs_pos.size(0) == m.voxel_obs_count.size(0)
        idx = m.indexer.cpu().numpy()
        occ = np.nonzero(idx != -1)[0]
        assert np.array_equal(occ, fx[f"f{f}.occ_cells"]) and np.array_equal(idx[occ], fx[f"f{f}.occ_slots"])       # bit-exact index
        pos = m.latent_vecs_pos.cpu().numpy()
        assert np.array_equal(pos[:nocc][fx[f"f{f}.occ_slots"]], fx[f"f{f}.occ_cells"]) and np.all(pos[nocc:] == -1)
        assert np.array_equal(m.voxel_obs_count.cpu().numpy()[:nocc], fx[f"f{f}.obs_count"])                       # exact
        lat = m.latent_vecs.cpu().numpy()
        assert close(lat[fx[f"f{f}.latent_rows"]], fx[f"f{f}.latent"], TOL)
        assert abs(lat[:nocc].astype(np.float64).sum() - float(fx[f"f{f}.latent_sum"])) < 1e-5 * np.abs(lat[:nocc]).sum()   # checksum over all rows
        assert np.array_equal(m.mesh_cache.updated_vec_id.cpu().numpy(), fx[f"f{f}.updated_vec_id"])
        st = m.last_integrate_stats
        assert st["n_kept"] == int(mask.sum()) and st["n_updated"] >= 0

    # get_sdf + autograd contract (map.py:559-579, tracker.py:183-194)
    q = _t(fx["q.xyz"], dev).requires_grad_(True)
    sdf, std, valid = m.get_sdf(q)
    assert np.array_equal(np.packbits(valid.cpu().numpy()), fx["q.valid"])
    assert close(sdf.detach().cpu().numpy(), fx["q.sdf"], TOL) and close(std.detach().cpu().numpy(), fx["q.std"], TOL)
    r = sdf / std.detach()
    grad = torch.autograd.grad(r, [q], grad_outputs=torch.ones_like(r))[0]
    assert grad.shape == q.shape and bool((grad[~valid] == 0).all())
    assert frac_off(grad[valid].cpu().numpy(), fx["q.grad"]) < 2e-3

    if "hg.H" in fx.files:                                    # compute_sdf_Hg (tracker.py:174-218)
        trk = SDFTracker(m, argparse.Namespace(sdf=dict(robust_kernel="huber", robust_k=5.0, subsample=0.5), rgb=None,
                                               iter_config=[{"n": 50, "type": [["sdf"]]}]))
        last = Isometry(q=Rotation(matrix=fx["hg.R_last"]), t=fx["hg.t_last"])
        delta = Isometry(q=Rotation(matrix=fx["hg.R_delta"]), t=fx["hg.t_delta"])
        H, g, E = trk.compute_sdf_Hg(0, last, delta, _t(fx["hg.obs"], dev), no_grad=False)
        assert H.shape == (6, 6) and H.dtype == np.float64 and np.allclose(H, H.T)
        # element-wise at the Gram scale (conftest.hg_errors), and far inside the old max-norm bound
        gram, strict = hg_errors(H, g, E, fx["hg.H"], fx["hg.g"], float(fx["hg.E"]))
        print(f"[{name}] H,g vs the executed reference: err/tol {gram:.3f} at the Gram scale, {strict:.2f} at |b| (cancelled off-diagonals)")
        assert gram <= 1.0
        assert np.abs(H - fx["hg.H"]).max() <= 1e-4 * np.abs(fx["hg.H"]).max()        # (the CPU restatement itself: 1e-4, tests/test_oracle_golden.py)
        assert np.abs(g - fx["hg.g"]).max() <= 1e-4 * np.abs(fx["hg.g"]).max()
        assert close(E, float(fx["hg.E"]), TOL)
        H2, g2, E2 = trk.compute_sdf_Hg(-1, last, delta, _t(fx["hg.obs"], dev), no_grad=True)
        assert H2 is None and g2 is None and close(E2, float(fx["hg.E_nograd"]), TOL)

    if "mesh.res" in fx.files:                                # mesh decode (map.py:624-687)
        r = int(fx["mesh.res"])
        focused, mapping, cs, cd, slots, cnt = m.mesh_cubes(r, fast=True, updated_vec_id=None)
        assert np.array_equal(focused.cpu().numpy(), fx["mesh.focused"])
        ref_map = fx["mesh.mapping"]
        assert np.array_equal(mapping.cpu().numpy()[:ref_map.shape[0]], ref_map) and bool((mapping[ref_map.shape[0]:] == -1).all())
        assert cs.shape == (int(fx["mesh.B"]), 2 * r, 2 * r, 2 * r)
        sel = fx["mesh.sel"]
        cs_n, cd_n = cs.cpu().numpy(), cd.cpu().numpy()
        # the |sdf|<0.05 re-evaluation set can differ by threshold flips on 1e-6 noise (SURVEY 7 "discrete decisions")
        assert frac_off(cs_n[sel], fx["mesh.sdf_sel"]) < 1e-3 and frac_off(cd_n[sel], fx["mesh.std_sel"]) < 1e-3
        assert abs(cs_n.astype(np.float64).sum() - float(fx["mesh.sdf_sum"])) < 0.5
        # MC on our own cubes: CUDA kernel vs the scalar restatement, identical inputs => identical triangle multiset
        _assert_mc_equal(m.indexer.view(m.n_xyz), focused, mapping, cs, cd, m.n_xyz, float(fx["mesh.max_std"]))
        mesh = m.extract_mesh(r, int(4e6), max_std=float(fx["mesh.max_std"]), no_cache=True)
        assert abs(mesh.triangles.shape[0] - int(fx["mesh.n_tri_oracle_mc"])) <= 64
        assert mesh.vertices.shape[0] == 3 * mesh.triangles.shape[0]


def _sorted_tris(tri, fid, std):
    key = np.concatenate([fid[:, None].astype(np.float64), tri.reshape(len(tri), 9).astype(np.float64)], 1)
    order = np.lexsort(key.T[::-1])
    return tri[order], fid[order], std[order]


def _assert_mc_equal(indexer, focused, mapping, cs, cd, n_xyz, max_std, max_tri=int(4e6)):
    from difusion_b200.system import ext
    from oracle import mc_oracle
    tri, fid, std = ext.marching_cubes_interp(indexer, focused, mapping, cs, cd, max_tri, n_xyz, max_std)
    o_tri, o_fid, o_std = mc_oracle.marching_cubes_interp(indexer.cpu().numpy(), focused.cpu().numpy(), mapping.cpu().numpy(),
                                                          cs.cpu().numpy(), cd.cpu().numpy(), max_tri, n_xyz, max_std)
    assert tri.shape[0] == o_tri.shape[0] > 0
    a = _sorted_tris(tri.cpu().numpy(), fid.cpu().numpy(), std.cpu().numpy())
    b = _sorted_tris(o_tri, o_fid, o_std)
    assert np.array_equal(a[1], b[1])
    assert np.abs(a[0] - b[0]).max() <= 1e-6 and np.abs(a[2] - b[2]).max() <= 1e-6
    return tri.shape[0]


@pytest.mark.parametrize("r", [4, 5, 2])
def test_marching_cubes_synthetic(dev, r):
    """system.ext.marching_cubes_interp on analytic cubes: parity with the scalar restatement, max_std filter, overflow."""
    from difusion_b200.system import ext
    from test_oracle_mc import _sphere_cubes
    n_xyz = [6, 5, 7]
    indexer, mapping, sdf, std = _sphere_cubes(n_xyz, r, (3.0, 2.5, 3.5), 1.9, drop=(2, 2, 3))
    rng = np.random.default_rng(0)
    std = (std + rng.uniform(0, 0.1, std.shape)).astype(np.float32)
    # leave some PLIVoxes out of the batch (mapping -1) and shuffle batch order
    B = sdf.shape[0]
    perm = rng.permutation(B)
    keep = perm[: int(B * 0.9)]
    mapping = np.full(B, -1, np.int32); mapping[keep] = np.arange(len(keep), dtype=np.int32)
    sdf, std = sdf[keep], std[keep]
    blocks = np.sort(rng.choice(B, int(B * 0.8), replace=False)).astype(np.int64)
    blocks = blocks[indexer.reshape(-1)[blocks] != -1]
    args = (_t(indexer, dev), _t(blocks, dev), _t(mapping, dev), _t(sdf, dev), _t(std, dev))
    n_all = _assert_mc_equal(*args, n_xyz, 10.0)
    n_f = _assert_mc_equal(*args, n_xyz, 0.15)
    assert 0 < n_f < n_all
    tri, fid, tstd = ext.marching_cubes_interp(*args, 50, n_xyz, 10.0)        # overflow: full untrimmed buffer (mc_interp_kernel.cu:375-379)
    assert tri.shape == (50, 3, 3) and fid.shape == (50,) and tstd.shape == (50, 3)
    with pytest.raises(RuntimeError):
        ext.marching_cubes_interp(args[0].cpu(), *args[1:], 50, n_xyz, 10.0)


def test_groupby_sum(dev):
    from difusion_b200.system import ext
    from difusion_b200.network import utility as net_util
    g = torch.Generator().manual_seed(0)
    v = torch.randn(5000, 29, generator=g)
    idx = torch.randint(0, 37, (5000,), generator=g)
    s, c = ext.groupby_sum(v.to(dev), idx.to(dev), 40)
    ref = torch.zeros(40, 29).index_add_(0, idx, v)
    assert close(s.cpu().numpy(), ref.numpy(), 1e-4)
    assert np.array_equal(c.cpu().numpy(), 29 * np.bincount(idx.numpy(), minlength=40))       # indexing.cu:70 counts per column
    assert c.dtype == torch.int32 and s.shape == (40, 29)
    r = net_util.groupby_reduce(idx.to(dev), v.to(dev), op="sum")
    assert close(r.cpu().numpy(), ref.numpy()[: int(idx.max()) + 1], 1e-4)


def test_edge_cases(model, dev):
    from difusion_b200.system.map import DenseIndexedMap
    from difusion_b200 import synthetic as S
    from oracle import dif_oracle as O
    sc = S.scene_S0()
    W = O.load_weights_npz(GOLDEN / "weights.npz")
    # (1) points exactly on cell faces and on the grid border; prune disabled -> returns None (map.py:372-373)
    args = sc.map_args(); args.prune_min_vox_obs = 0; args.ignore_count_th = 0.0
    m = DenseIndexedMap(model, args, 29, dev)
    o = O.OracleMap(W, args)
    bm = np.asarray(args.bound_min, np.float32)
    ijk = np.array([[1, 1, 1], [1, 2, 3], [31, 31, 31], [32, 32, 32], [5, 5, 5], [0.5, 0.5, 0.5], [31.99, 0.01, 16]], np.float32)
    pts = (bm + ijk * np.float32(0.1)).astype(np.float32)
    pts = np.concatenate([pts, pts + np.float32(1e-6), pts[:5] - np.float32(1e-6)])
    pts = pts[np.all(np.ceil(o._normalize(pts)) - 1 >= 0, 1) & np.all(np.ceil(o._normalize(pts)) - 1 <= 31, 1)]
    nrm = np.tile(np.array([[0, 0, 1.0]], np.float32), (len(pts), 1))
    assert m.integrate_keyframe(_t(pts, dev), _t(nrm, dev)) is None and o.integrate_keyframe(pts, nrm) is None
    assert m.n_occupied == o.n_occupied and np.array_equal(m.indexer.cpu().numpy(), o.indexer)
    assert np.array_equal(m.voxel_obs_count.cpu().numpy(), o.voxel_obs_count)
    assert close(m.latent_vecs.cpu().numpy(), o.latent_vecs, TOL)
    # (2) everything pruned: nothing allocated, mask all False
    m2 = DenseIndexedMap(model, sc.map_args(), 29, dev)
    mask = m2.integrate_keyframe(_t(pts, dev), _t(nrm, dev))
    assert not bool(mask.any()) and m2.n_occupied == 0 and bool((m2.indexer == -1).all()) and m2.latent_vecs.size(0) == 1
    # (3) empty input
    e = torch.zeros((0, 3), device=dev)
    mask = m2.integrate_keyframe(e, e)
    assert mask.numel() == 0 and m2.n_occupied == 0
    # (4) out-of-bounds points are dropped and reported
    far = _t(np.array([[100.0, 0, 0]], np.float32), dev)
    m2.integrate_keyframe(far, far)
    assert m2.n_occupied == 0 and m2.last_integrate_stats["flags"] & 1 and m2.n_frames_with_dropped_points == 1
    # (5) get_sdf on an empty map asserts like the reference (utility.py:84-85)
    with pytest.raises(AssertionError):
        m2.get_sdf(_t(pts, dev))
    # (6) device mismatch assert (map.py:353)
    with pytest.raises(AssertionError):
        m2.integrate_keyframe(torch.zeros(3, 3), torch.zeros(3, 3))


def test_capacity_growth_and_save_load(golden, model, dev, tmp_path):
    """tiny physical capacity forces the doubling path; save/load round-trips the reference's cold_vars dict (map.py:239-249)."""
    from difusion_b200.system.map import DenseIndexedMap
    fx = golden["s0_map"]
    m = DenseIndexedMap(model, fixture_args(fx), 29, dev, initial_capacity=4)
    for f in range(int(fx["n_frames"])):
        m.integrate_keyframe(_t(fx[f"f{f}.xyz"], dev), _t(fx[f"f{f}.normal"], dev))
    f = int(fx["n_frames"]) - 1
    assert m.n_occupied == int(fx[f"f{f}.n_occupied"])
    assert np.array_equal(m.voxel_obs_count.cpu().numpy()[:m.n_occupied], fx[f"f{f}.obs_count"])
    assert close(m.latent_vecs.cpu().numpy()[fx[f"f{f}.latent_rows"]], fx[f"f{f}.latent"], TOL)
    p = tmp_path / "map.pt"
    m.save(p)
    cv = torch.load(p)
    assert set(cv) == {"n_occupied", "indexer", "latent_vecs", "latent_vecs_pos", "voxel_obs_count", "voxel_optimized"}
    assert cv["latent_vecs"].shape == (int(fx[f"f{f}.capacity"]), 29) and cv["voxel_optimized"].dtype == torch.bool
    m2 = DenseIndexedMap(model, fixture_args(fx), 29, dev)
    m2.load(p)
    assert m2.n_occupied == m.n_occupied and torch.equal(m2.indexer, m.indexer) and torch.equal(m2.latent_vecs, m.latent_vecs)
    q = _t(fx["q.xyz"], dev)
    a, b = m.get_sdf(q), m2.get_sdf(q)
    assert torch.equal(a[0], b[0]) and torch.equal(a[2], b[2])


def test_large_batch_properties(model, dev):
    """At BASELINE config-3 sizes (2^20 here, 2^22 in bench.py) the oracle is too slow: use size-independent properties."""
    from difusion_b200.network import utility as net_util
    from oracle import dif_oracle as O
    W = O.load_weights_npz(GOLDEN / "weights.npz")
    g = torch.Generator().manual_seed(5)
    n = 1 << 20
    table = torch.randn(2048, 29, generator=g) * 0.2
    rows = torch.randint(0, 2048, (n,), generator=g)
    xyz = torch.rand(n, 3, generator=g) * 2 - 1
    lat = table[rows]
    sdf, std = net_util.forward_model(model.decoder, latent_input=lat.to(dev), xyz_input=xyz.to(dev))
    sdf, std = sdf.cpu()[:, 0], std.cpu()[:, 0]
    # (a) a seeded sub-sample against the oracle
    pick = torch.randperm(n, generator=g)[:8192]
    o_sdf, o_std = O.decoder_forward(W.dec, lat[pick], xyz[pick])
    assert close(sdf[pick].numpy(), o_sdf.numpy(), TOL) and close(std[pick].numpy(), o_std.numpy(), TOL)
    # (b) permutation equivariance: decode(perm(x)) == perm(decode(x)) bit for bit (no cross-sample coupling)
    perm = torch.randperm(n, generator=g)
    sdf_p, _ = net_util.forward_model(model.decoder, latent_input=lat[perm].to(dev), xyz_input=xyz[perm].to(dev))
    assert torch.equal(sdf_p.cpu()[:, 0], sdf[perm])
    # (c) ranges: tanh-bounded sdf, std > 0.05 (di_decoder.py:68,84)
    assert float(sdf.abs().max()) <= 1.0 and float(std.min()) > 0.05
    # (d) analytic gradient vs central finite differences on the CUDA forward
    x0 = xyz[:4096].to(dev).requires_grad_(True)
    s0, _ = net_util.forward_model(model.decoder, latent_input=lat[:4096].to(dev), xyz_input=x0, no_detach=True)
    ga = torch.autograd.grad(s0.sum(), x0)[0].cpu()
    h = 1e-3
    for c in range(3):
        d = torch.zeros(1, 3); d[0, c] = h
        sp, _ = net_util.forward_model(model.decoder, latent_input=lat[:4096].to(dev), xyz_input=(xyz[:4096] + d).to(dev))
        sm, _ = net_util.forward_model(model.decoder, latent_input=lat[:4096].to(dev), xyz_input=(xyz[:4096] - d).to(dev))
        fd = ((sp - sm) / (2 * h)).cpu()[:, 0]
        assert float((fd - ga[:, c]).abs().median()) < 2e-3


def test_decoder_config3_full_size(model, dev):
    """BASELINE config 3 at its full size (2^22 samples): a seeded 32 k sub-sample against the oracle (chunked), the head and the
    ragged tail of the batch bit-identical to the same samples decoded in a small batch (tile position must not matter)."""
    from difusion_b200.network import utility as net_util
    from oracle import dif_oracle as O
    W = O.load_weights_npz(GOLDEN / "weights.npz")
    g = torch.Generator().manual_seed(11)
    n = (1 << 22) + 77                                          # ragged last tile
    table = torch.randn(23000, 29, generator=g) * 0.2
    rows = torch.randint(0, 23000, (n,), generator=g)
    xyz = torch.rand(n, 3, generator=g) * 2 - 1
    lat_d, xyz_d = table.to(dev)[rows.to(dev)], xyz.to(dev)
    sdf, std = net_util.forward_model(model.decoder, latent_input=lat_d, xyz_input=xyz_d)
    assert sdf.shape == (n, 1) and bool(torch.isfinite(sdf).all()) and bool(torch.isfinite(std).all())
    pick = torch.randperm(n, generator=g)[:32768]
    pick[:128] = torch.arange(n - 128, n)                       # the ragged tail is in the sample
    for c in range(0, pick.numel(), 8192):
        pc_ = pick[c:c + 8192]
        o_sdf, o_std = O.decoder_forward(W.dec, table[rows[pc_]], xyz[pc_])
        assert close(sdf.cpu()[pc_, 0].numpy(), o_sdf.numpy(), TOL) and close(std.cpu()[pc_, 0].numpy(), o_std.numpy(), TOL)
    pd = pick.to(dev)
    s_small, u_small = net_util.forward_model(model.decoder, latent_input=lat_d[pd], xyz_input=xyz_d[pd])
    assert torch.equal(s_small, sdf[pd]) and torch.equal(u_small, std[pd])


def test_large_map_properties(model, dev):
    """BASELINE config-5 scale: a 1000 x 1000 x 40 = 40 M-cell dense index (320 MB), S1-sized views into a height field.
    The oracle cannot hold this in reasonable time, so the integer state is checked against an independent torch
    formulation of the allocation rule (map.py:366-387) and the floating-point state through invariants."""
    import argparse
    from difusion_b200.system.map import DenseIndexedMap
    vs = 0.05
    args = argparse.Namespace(bound_min=[0.0, 0.0, 0.0], bound_max=[50.0, 50.0, 2.0], voxel_size=vs, prune_min_vox_obs=2,
                              ignore_count_th=4.0, encoder_count_th=600.0)
    m = DenseIndexedMap(model, args, 29, dev, initial_capacity=1 << 20)
    assert m.n_xyz == [1000, 1000, 40] and m.indexer.numel() == 40_000_000
    g = torch.Generator().manual_seed(21)
    occupied = torch.zeros(40_000_000, dtype=torch.bool, device=dev)
    total = 0
    for f in range(6):
        # a 6 m x 6 m patch of the height field z = 1 + 0.4 sin(x) cos(0.7 y), ~30k points, patches overlap between frames
        x0, y0 = 5.0 + 4.0 * f, 43.0 - 6.0 * f
        xy = torch.rand(30_000, 2, generator=g) * 6.0 + torch.tensor([x0, y0])
        z = 1.0 + 0.4 * torch.sin(xy[:, 0]) * torch.cos(0.7 * xy[:, 1])
        pts = torch.cat([xy, z[:, None]], 1).float().to(dev)
        nrm = torch.tensor([[0.0, 0.0, 1.0]]).repeat(pts.size(0), 1).to(dev)
        n_before = m.n_occupied
        mask = m.integrate_keyframe(pts, nrm)
        # independent restatement of voxelise + prune + allocate with sort-based torch ops
        # (a 0-dim CUDA divisor forces a true fp32 division like torch-CPU, the oracle's arithmetic; with a Python-float
        #  divisor torch-CUDA multiplies by the fp32 reciprocal instead, which moves points that sit on a cell face)
        ijk = (torch.ceil((pts - torch.tensor(args.bound_min, device=dev)) / torch.tensor(vs, device=dev)) - 1).long()
        lin = ijk[:, 2] + 40 * ijk[:, 1] + 40_000 * ijk[:, 0]
        uq, inv, cnt = torch.unique(lin, return_inverse=True, return_counts=True)
        keep = cnt[inv] > 2
        assert torch.equal(mask, keep)
        kept = ijk[keep]
        kept = kept[~occupied[lin[keep]]]                                 # E0: EMPTY cells hit by a kept point (map.py:380-383)
        nb = torch.tensor([[0, 0, 0], [1, 0, 0], [-1, 0, 0], [0, 1, 0], [0, -1, 0], [0, 0, 1], [0, 0, -1]], device=dev)
        cand = (kept[:, None, :] + nb[None]).reshape(-1, 3)
        hi = torch.tensor([999, 999, 39], device=dev)
        cand = torch.minimum(torch.maximum(cand, torch.zeros_like(hi)), hi)
        cl = torch.unique(cand[:, 2] + 40 * cand[:, 1] + 40_000 * cand[:, 0])
        new = cl[~occupied[cl]]                                           # ascending linear id == slot order
        occupied[new] = True
        assert m.n_occupied == n_before + new.numel()
        assert torch.equal(m.latent_vecs_pos[n_before:m.n_occupied], new)
        assert torch.equal(m.indexer[new], torch.arange(n_before, m.n_occupied, device=dev))
        total += int(keep.sum())
    n = m.n_occupied
    assert int((m.indexer >= 0).sum()) == n == int(occupied.sum())
    assert torch.equal(m.indexer[m.latent_vecs_pos[:n]], torch.arange(n, device=dev))          # indexer and pos are inverse maps
    obs = m.voxel_obs_count[:n]
    lat = m.latent_vecs[:n]
    assert bool(torch.isfinite(lat).all()) and float(obs.min()) >= 0 and bool((obs == obs.round()).all())
    assert bool((lat[obs == 0] == 0).all()) and bool((lat[obs > 0].abs().sum(1) > 0).all())
    # every kept point contributes to at most 8 cells, at least its own
    assert total <= float(obs.sum()) <= 8 * total
    # decode through the big index: samples in observed cells are valid, far-away ones are not
    q = torch.cat([pts[:4096], pts[:16] + torch.tensor([0.0, 0.0, 0.9], device=dev)])
    sdf, std, valid = m.get_sdf(q)
    assert int(valid[:4096].sum()) > 2000 and not bool(valid[4096:].any()) and float(sdf.abs().max()) <= 1.0 and float(std.min()) > 0.05


def test_tensor_core_decoder_matches_fp32_path(golden, model, dev):
    """tcgen05 forward (3-pass fp16 split, fp32 TMEM accumulation) vs the exact-fp32 SIMT kernel and vs the reference fixture.
    Forward-only launches of >= 1024 samples take the tensor-core kernel; a launch that also asks for d/dxyz takes the SIMT kernel."""
    from difusion_b200.network import utility as net_util
    fx = golden["decoder_kat"]
    for n in (4096, 1024, 3000, 1025):                       # full tiles, ragged tail, single CTA with one / two slots
        lat, xyz = _t(fx["latent"][:n], dev), _t(fx["xyz"][:n], dev)
        sdf_tc, std_tc = net_util.forward_model(model.decoder, latent_input=lat, xyz_input=xyz)
        sdf_32, std_32 = net_util.forward_model(model.decoder, latent_input=lat, xyz_input=xyz.clone().requires_grad_(True), no_detach=True)
        a, b = sdf_tc.cpu().numpy()[:, 0], sdf_32.detach().cpu().numpy()[:, 0]
        assert np.abs(a - b).max() < 2e-5, np.abs(a - b).max()
        assert np.abs(std_tc.cpu().numpy() - std_32.detach().cpu().numpy()).max() < 2e-5
        assert close(a, fx["sdf"][:n], TOL) and close(std_tc.cpu().numpy()[:, 0], fx["std"][:n], TOL)
    # many tiles per CTA (persistent loop, both TMEM slots, phase bits wrapping) + large-magnitude latents
    g = torch.Generator().manual_seed(11)
    n = 148 * 128 * 5 + 77
    lat = torch.randn(n, 29, generator=g) * 0.5
    xyz = torch.rand(n, 3, generator=g) * 2 - 1
    sdf_tc, std_tc = net_util.forward_model(model.decoder, latent_input=lat.to(dev), xyz_input=xyz.to(dev))
    from oracle import dif_oracle as O
    W = O.load_weights_npz(GOLDEN / "weights.npz")
    pick = torch.cat([torch.arange(0, 4096), torch.arange(n - 4096, n), torch.randperm(n, generator=g)[:8192]])
    o_sdf, o_std = O.decoder_forward(W.dec, lat[pick], xyz[pick])
    assert close(sdf_tc.cpu()[pick, 0].numpy(), o_sdf.numpy(), TOL) and close(std_tc.cpu()[pick, 0].numpy(), o_std.numpy(), TOL)


def test_tensor_core_icp_matches_fp32_path(golden, model, dev, monkeypatch):
    """compute_sdf_Hg through the tcgen05 forward+backward kernel vs the exact-fp32 SIMT kernel (DIF_ICP_PATH=simt) on the
    same map and points, with and without gradients, including a pose that leaves many points outside observed PLIVoxes."""
    import argparse
    from difusion_b200.system.map import DenseIndexedMap
    from difusion_b200.system.tracker import SDFTracker
    from difusion_b200.utils.motion_util import Isometry, Rotation
    fx = golden["s1_map"]
    m = DenseIndexedMap(model, fixture_args(fx), 29, dev)
    for f in range(int(fx["n_frames"])):
        m.integrate_keyframe(_t(fx[f"f{f}.xyz"], dev), _t(fx[f"f{f}.normal"], dev))
    trk = SDFTracker(m, argparse.Namespace(sdf=dict(robust_kernel="huber", robust_k=5.0, subsample=0.5), rgb=None,
                                           iter_config=[{"n": 2, "type": [["sdf"]]}]))
    last = Isometry(q=Rotation(matrix=fx["hg.R_last"]), t=fx["hg.t_last"])
    obs = _t(fx["hg.obs"], dev)
    for xi in ([0, 0, 0, 0, 0, 0], [0.004, -0.003, 0.002, 0.003, -0.002, 0.001], [0.3, 0.1, -0.2, 0.05, 0.02, -0.04]):
        delta = Isometry.from_twist(np.asarray(xi, float))
        monkeypatch.delenv("DIF_ICP_PATH", raising=False)
        H, g, E = trk.compute_sdf_Hg(0, last, delta, obs, no_grad=False)
        _, _, E_ng = trk.compute_sdf_Hg(-1, last, delta, obs, no_grad=True)
        n_tc = float(m.icp_linearize(obs, last.q.rotation_matrix, last.t, delta.q.rotation_matrix, delta.t)[43])
        monkeypatch.setenv("DIF_ICP_PATH", "simt")
        H2, g2, E2 = trk.compute_sdf_Hg(0, last, delta, obs, no_grad=False)
        n_simt = float(m.icp_linearize(obs, last.q.rotation_matrix, last.t, delta.q.rotation_matrix, delta.t)[43])
        monkeypatch.delenv("DIF_ICP_PATH", raising=False)
        assert n_tc == n_simt > 1000                              # identical valid sets (integer lookup path)
        assert hg_errors(H, g, E, H2, g2, E2)[0] <= 1.0           # element-wise, Gram scale (conftest.hg_errors)
        assert np.abs(H - H2).max() <= 5e-5 * np.abs(H2).max() and np.abs(g - g2).max() <= 5e-5 * np.abs(g2).max()
        assert close(E, E2, 1e-5) and close(E_ng, E2, 1e-5)
    # ragged sizes / single tile / one slot only
    for n in (2048, 2049, 4000):
        delta = Isometry.from_twist(np.zeros(6))
        o = m.icp_linearize(obs[:n], last.q.rotation_matrix, last.t, delta.q.rotation_matrix, delta.t).cpu().numpy()
        monkeypatch.setenv("DIF_ICP_PATH", "simt")
        o2 = m.icp_linearize(obs[:n], last.q.rotation_matrix, last.t, delta.q.rotation_matrix, delta.t).cpu().numpy()
        monkeypatch.delenv("DIF_ICP_PATH", raising=False)
        assert o[43] == o2[43] and np.abs(o[:36] - o2[:36]).max() <= 5e-5 * np.abs(o2[:36]).max() and close(o[42], o2[42], 1e-5)
    # many tiles per slot (producer warps, early F0 behind B0, barrier phases over 16 iterations), with and without gradients
    big = obs.repeat(16, 1).contiguous()
    delta = Isometry.from_twist(np.asarray([0.004, -0.003, 0.002, 0.003, -0.002, 0.001]))
    n_one = float(m.icp_linearize(obs, last.q.rotation_matrix, last.t, delta.q.rotation_matrix, delta.t)[43])
    for grad in (True, False):
        o = m.icp_linearize(big, last.q.rotation_matrix, last.t, delta.q.rotation_matrix, delta.t, 5.0, grad).cpu().numpy()
        monkeypatch.setenv("DIF_ICP_PATH", "simt")
        o2 = m.icp_linearize(big, last.q.rotation_matrix, last.t, delta.q.rotation_matrix, delta.t, 5.0, grad).cpu().numpy()
        monkeypatch.delenv("DIF_ICP_PATH", raising=False)
        assert o[43] == o2[43] == 16 * n_one and close(o[42], o2[42], 1e-5)
        if grad:
            assert hg_errors(o[:36].reshape(6, 6), o[36:42], o[42], o2[:36].reshape(6, 6), o2[36:42], o2[42])[0] <= 1.0
    # the first pipeline (DIF_ICP_V=1, kept for A/B timing) still agrees
    monkeypatch.setenv("DIF_ICP_V", "1")
    o1 = m.icp_linearize(obs, last.q.rotation_matrix, last.t, delta.q.rotation_matrix, delta.t).cpu().numpy()
    monkeypatch.delenv("DIF_ICP_V", raising=False)
    o = m.icp_linearize(obs, last.q.rotation_matrix, last.t, delta.q.rotation_matrix, delta.t).cpu().numpy()
    assert o1[43] == o[43] and np.abs(o1[:42] - o[:42]).max() <= 5e-5 * np.abs(o[:42]).max()
    # the tensor-core kernel reduces through per-CTA partials in a fixed order: results are bit-reproducible run to run
    delta = Isometry.from_twist(np.asarray([0.004, -0.003, 0.002, 0.003, -0.002, 0.001]))
    runs = [m.icp_linearize(obs, last.q.rotation_matrix, last.t, delta.q.rotation_matrix, delta.t).cpu().numpy() for _ in range(4)]
    assert all(np.array_equal(runs[0], r) for r in runs[1:])
    assert np.array_equal(runs[0][:36].reshape(6, 6), runs[0][:36].reshape(6, 6).T)


def test_tensor_core_encoder_matches_fp32_path(golden, model, dev, monkeypatch):
    """dif_encode / dif_integrate through the tcgen05 encoder vs the exact-fp32 SIMT kernels (DIF_ENCODE_PATH=simt)."""
    from difusion_b200.system.map import DenseIndexedMap
    fx = golden["encoder_kat"]
    for n in (4096, 1024, 1025, 3000):
        x = _t(fx["xyzn"][:n], dev)
        monkeypatch.delenv("DIF_ENCODE_PATH", raising=False)
        tc = model.encoder(x).cpu().numpy()
        monkeypatch.setenv("DIF_ENCODE_PATH", "simt")
        ref = model.encoder(x).cpu().numpy()
        monkeypatch.delenv("DIF_ENCODE_PATH", raising=False)
        assert np.abs(tc - ref).max() < 2e-5, np.abs(tc - ref).max()
        assert close(tc, fx["latent"][:n], TOL)
    fm = golden["s1_map"]
    maps = {}
    for path in ("tc", "simt"):
        if path == "simt":
            monkeypatch.setenv("DIF_ENCODE_PATH", "simt")
        m = DenseIndexedMap(model, fixture_args(fm), 29, dev)
        for f in range(int(fm["n_frames"])):
            m.integrate_keyframe(_t(fm[f"f{f}.xyz"], dev), _t(fm[f"f{f}.normal"], dev))
        maps[path] = (m.n_occupied, m.indexer.cpu(), m.voxel_obs_count.cpu(), m.latent_vecs.cpu())
        monkeypatch.delenv("DIF_ENCODE_PATH", raising=False)
    a, b = maps["tc"], maps["simt"]
    assert a[0] == b[0] and torch.equal(a[1], b[1]) and torch.equal(a[2], b[2])
    assert float((a[3] - b[3]).abs().max()) < 2e-5


def test_mesh_scene_against_oracle(model, dev):
    """BASELINE config 4 at reduced size (sphere R=0.9 m, 5 cm PLIVoxes, voxel_resolution 5 = 1 cm): select + lattice decode +
    trilinear/select + marching cubes vs the oracle restatement of map.py:624-702 on the same map."""
    from difusion_b200 import synthetic as S
    from difusion_b200.system.map import DenseIndexedMap
    from oracle import dif_oracle as O
    R, n_pts = 0.9, 200_000
    sc = S.Scene("S2s", [-1.1] * 3, [1.1] * 3, 0.05, 2, 4.0)
    i = np.arange(n_pts) + 0.5
    phi = np.arccos(1 - 2 * i / n_pts); th = np.pi * (1 + 5 ** 0.5) * i
    d = np.stack([np.cos(th) * np.sin(phi), np.sin(th) * np.sin(phi), np.cos(phi)], 1)
    pts, nrm = (R * d).astype(np.float32), (-d).astype(np.float32)
    W = O.load_weights_npz(GOLDEN / "weights.npz")
    m = DenseIndexedMap(model, sc.map_args(), 29, dev)
    o = O.OracleMap(W, sc.map_args())
    for c in range(2):
        sl = slice(c * n_pts // 2, (c + 1) * n_pts // 2)
        m.integrate_keyframe(_t(pts[sl], dev), _t(nrm[sl], dev))
        o.integrate_keyframe(pts[sl], nrm[sl])
    assert m.n_occupied == o.n_occupied and np.array_equal(m.indexer.cpu().numpy(), o.indexer)
    assert np.array_equal(m.voxel_obs_count.cpu().numpy(), o.voxel_obs_count)
    assert close(m.latent_vecs.cpu().numpy(), o.latent_vecs, TOL)
    for fast in (True, False):
        focused, mapping, cs, cd, slots, cnt = m.mesh_cubes(5, fast=fast, updated_vec_id=None)
        o_foc, o_map, o_hs, o_hd, o_occ = o.mesh_cubes(5, fast=fast, no_cache=True)
        assert np.array_equal(focused.cpu().numpy(), o_foc) and np.array_equal(slots.cpu().numpy(), o_occ)
        assert np.array_equal(mapping.cpu().numpy()[:o_map.shape[0]], o_map)
        # threshold flips of the |sdf|<0.05 re-evaluation set are allowed on a tiny fraction of samples
        assert frac_off(cs.cpu().numpy(), o_hs) < 2e-4 and frac_off(cd.cpu().numpy(), o_hd) < 2e-4
        if fast:
            assert abs(int(cnt[1]) - o.last_stats["n_high"]) <= max(20, o.last_stats["n_high"] // 2000)
        n_tri = _assert_mc_equal(m.indexer.view(m.n_xyz), focused, mapping, cs, cd, m.n_xyz, 0.15, max_tri=int(6e6))
        assert n_tri > 10000
    # incremental extraction: only PLIVoxes touched since the last extraction are re-meshed (map.py:610-622, 703-714)
    mesh_all = m.extract_mesh(5, int(6e6), max_std=0.15, no_cache=True)
    extra = (pts[:2000] * np.float32(1.0)).astype(np.float32)
    m.integrate_keyframe(_t(extra, dev), _t(nrm[:2000], dev))
    upd = m.mesh_cache.updated_vec_id
    assert 0 < upd.numel() < m.n_occupied
    mesh_inc = m.extract_mesh(5, int(6e6), max_std=0.15)
    assert mesh_inc.triangles.shape[0] > 0 and m.mesh_cache.updated_vec_id.numel() == 0
    assert abs(mesh_inc.triangles.shape[0] - mesh_all.triangles.shape[0]) < 0.2 * mesh_all.triangles.shape[0]


def test_device_mesh_cache_merge_matches_reference_host_merge(golden, model, dev):
    """dif_mesh_cache_merge vs the reference's host merge (map.py:698-714).  The keep masks come from the reference's
    own `_get_valid_idx` (:20-26) EXECUTED on these id streams (tests/golden/make_golden_merge.py -> ref_host_merge.npz); rows must be
    bit-exact, in the reference order (kept cached rows, then new rows), with the world transform of :698."""
    import ctypes
    from difusion_b200 import _lib
    from oracle import dif_oracle as O
    L = _lib.lib()
    fx = golden["ref_host_merge"]
    rng = np.random.default_rng(11)
    n_cells, vs, bmin = 50_000, np.float32(0.05), np.array([-1.5, 0.25, 2.0], np.float32)
    persist = torch.zeros(L.dif_mesh_cache_scratch_bytes(n_cells, 1 << 16), dtype=torch.uint8, device=dev)
    cache = None
    for step in range(5):
        fid = fx[f"s{step}.new_ids"]
        n_new = fid.shape[0]
        tri = rng.uniform(0, 40, (n_new, 3, 3)).astype(np.float32)
        std = rng.uniform(0, 0.15, (n_new, 3)).astype(np.float32)
        world = tri * vs + bmin                                                # map.py:698 (two rounded fp32 ops)
        if cache is None:
            exp = (world, fid, std)
        else:
            keep = np.unpackbits(fx[f"s{step}.keep"])[:cache[1].shape[0]].astype(bool)     # the executed reference's mask
            assert np.array_equal(keep, O.host_cache_keep_mask(cache[1], fid))
            exp = tuple(np.concatenate([c[keep], n], 0) for c, n in zip(cache, (world, fid, std)))
        assert exp[1].shape[0] == int(fx[f"s{step}.n_cache_after"])
        n_cache = 0 if cache is None else cache[0].shape[0]
        d_cache = [None] * 3 if cache is None else [_t(c, dev) for c in cache]
        d_new = [_t(tri, dev), _t(fid, dev), _t(std, dev)]
        o_tri = torch.empty((n_cache + n_new, 3, 3), device=dev); o_id = torch.empty(n_cache + n_new, dtype=torch.long, device=dev)
        o_std = torch.empty((n_cache + n_new, 3), device=dev); totals = torch.zeros(2, dtype=torch.long, device=dev)
        _lib.check(L.dif_mesh_cache_merge(_lib.ptr(d_cache[0]), _lib.ptr(d_cache[1]), _lib.ptr(d_cache[2]), n_cache,
                                          d_new[0].data_ptr(), d_new[1].data_ptr(), d_new[2].data_ptr(), n_new, float(vs),
                                          (ctypes.c_float * 3)(*bmin.tolist()), n_cells, o_tri.data_ptr(), o_id.data_ptr(), o_std.data_ptr(),
                                          totals.data_ptr(), persist.data_ptr(), persist.numel(), _lib.stream_ptr(dev)), "merge")
        kept, total = totals.tolist()
        assert total == exp[0].shape[0] and kept == total - n_new, (step, kept, total, exp[0].shape)
        assert np.array_equal(o_id[:total].cpu().numpy(), exp[1])
        assert np.array_equal(o_tri[:total].cpu().numpy().view(np.uint32), exp[0].view(np.uint32))
        assert np.array_equal(o_std[:total].cpu().numpy().view(np.uint32), exp[2].view(np.uint32))
        assert int(persist[:n_cells].sum()) == 0                               # the flag plane cleaned itself
        cache = exp
    # and through extract_mesh (lazy download, cache bookkeeping)
    from difusion_b200 import synthetic as S
    from difusion_b200.system.map import DenseIndexedMap
    sc = S.scene_S0()
    m = DenseIndexedMap(model, sc.map_args(), 29, dev)
    for yaw in (0.0, 0.15):
        R, t = S.yaw_pose(yaw); pc, nc = S.frame_points(sc, R, t); xw, nw = S.to_world(pc, nc, R, t)
        m.integrate_keyframe(_t(xw, dev), _t(nw, dev))
        upd_ids = m.latent_vecs_pos[m.mesh_cache.updated_vec_id].cpu().numpy()
        inc = m.extract_mesh(4, int(2e6), max_std=0.15)
    assert inc.n_triangles == m.mesh_cache.d_vertices.size(0) == m.mesh_cache.vertices.shape[0] > 1000
    assert inc.vertices.shape == (3 * inc.n_triangles, 3) and inc.triangles.shape == (inc.n_triangles, 3)
    v_inc, id_inc = m.mesh_cache.vertices.reshape(-1, 9).copy(), m.mesh_cache.vertices_flatten_id.copy()
    full = m.extract_mesh(4, int(2e6), max_std=0.15, no_cache=True)
    v_full, id_full = m.mesh_cache.vertices.reshape(-1, 9), m.mesh_cache.vertices_flatten_id
    assert 0.9 * full.n_triangles < inc.n_triangles < 1.1 * full.n_triangles
    # PLIVoxes re-meshed incrementally see only their 6 face neighbours in the decode batch (map.py:628-631), a full extraction
    # sees all 26, so rows may differ where a diagonal neighbour contributes; most re-meshed triangles are identical bit for bit
    a = np.unique(v_inc[np.isin(id_inc, upd_ids)], axis=0)
    b = set(map(bytes, np.unique(v_full[np.isin(id_full, upd_ids)], axis=0)))
    assert a.shape[0] > 500 and sum(bytes(r) in b for r in a) > 0.8 * a.shape[0]
    assert set(np.unique(id_inc[np.isin(id_inc, upd_ids)]).tolist()) <= set(upd_ids.tolist())


def test_config5_terrain_against_oracle(model, dev, oracle_weights):
    """BASELINE configs[4] / SURVEY 8(d) scene S3 (the 5 cm height field the sharded stream runs on), built with bulk integrate_keyframe
    calls exactly as bench.py's config-5 extra builds it, against the CPU oracle: integer state bit-exact, latents at the stated bar.
    Default: the FULL map (50 m x 50 m, 1000 x 1000 x 40 grid, 6.15 M points, 2.98 M allocated / 2.49 M observed PLIVoxes; ~35 s, nearly
    all of it the CPU oracle; measured max |latent - oracle| 1.1e-5); DIF_TEST_S3_EXTENT=25 runs a quarter of the area in ~10 s."""
    import os
    from difusion_b200 import synthetic as S
    from difusion_b200.system.map import DenseIndexedMap
    from oracle import dif_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    ext = float(os.environ.get("DIF_TEST_S3_EXTENT", "50"))
    sc = S.scene_S3(extent=ext)
    m = DenseIndexedMap(model, sc.map_args(), 29, dev, initial_capacity=1 << 20)
    o = O.OracleMap(oracle_weights, sc.map_args())
    for p, n in S.s3_terrain_points(extent=ext, rows_per_batch=200):
        mask = m.integrate_keyframe(torch.from_numpy(p).to(dev), torch.from_numpy(n).to(dev))
        o_mask = o.integrate_keyframe(p, n)
        assert np.array_equal(mask.cpu().numpy(), o_mask)
    n_occ = o.n_occupied
    assert m.n_occupied == n_occ and n_occ > 700_000 * (ext / 25.0) ** 2
    assert np.array_equal(m.indexer.cpu().numpy(), o.indexer)
    assert np.array_equal(m.latent_vecs_pos.cpu().numpy()[:n_occ], o.latent_vecs_pos[:n_occ])
    assert np.array_equal(m.voxel_obs_count.cpu().numpy()[:n_occ], o.voxel_obs_count[:n_occ])
    lat, ref = m.latent_vecs.cpu().numpy()[:n_occ], o.latent_vecs[:n_occ]
    assert close(lat, ref), float(np.abs(lat - ref).max())
    print(f"[config 5] extent {ext} m: {n_occ} PLIVoxes, {int((o.voxel_obs_count[:n_occ] > 0).sum())} observed; max |latent - oracle| {np.abs(lat - ref).max():.2e}")


def test_icp_tukey_kernel(golden, model, dev, monkeypatch, oracle_weights):
    """The sdf term with the Tukey robust kernel (tracker.py:66-69; the shipped config uses Huber): tensor-core path vs exact-fp32 SIMT
    path vs the CPU oracle, on the S0 fixture map (robust_k small enough that a third of the residuals fall outside the kernel)."""
    import argparse
    from difusion_b200.system.map import DenseIndexedMap
    from difusion_b200.system.tracker import SDFTracker
    from difusion_b200.utils.motion_util import Isometry, Rotation
    from oracle import dif_oracle as O
    fx = golden["s0_map"]
    m = DenseIndexedMap(model, fixture_args(fx), 29, dev)
    o = O.OracleMap(oracle_weights, fixture_args(fx))
    for f in range(int(fx["n_frames"])):
        m.integrate_keyframe(_t(fx[f"f{f}.xyz"], dev), _t(fx[f"f{f}.normal"], dev))
        o.integrate_keyframe(fx[f"f{f}.xyz"], fx[f"f{f}.normal"])
    k = 1.0
    trk = SDFTracker(m, argparse.Namespace(sdf=dict(robust_kernel="tukey", robust_k=k, subsample=0.5), rgb=None, iter_config=[]))
    last = Isometry(q=Rotation(matrix=fx["hg.R_last"]), t=fx["hg.t_last"])
    delta = Isometry.from_twist(np.asarray([0.004, -0.003, 0.002, 0.003, -0.002, 0.001]))
    obs = _t(fx["hg.obs"], dev)
    assert obs.size(0) >= 2048                                   # tensor-core path
    H, g, E = trk.compute_sdf_Hg(0, last, delta, obs, no_grad=False)
    monkeypatch.setenv("DIF_ICP_PATH", "simt")
    H2, g2, E2 = trk.compute_sdf_Hg(0, last, delta, obs, no_grad=False)
    monkeypatch.delenv("DIF_ICP_PATH", raising=False)
    oH, og, oE = O.compute_sdf_Hg(o, fx["hg.R_last"], fx["hg.t_last"], delta.q.rotation_matrix, delta.t, fx["hg.obs"], k, robust_kernel="tukey")
    hH, hg_, hE = O.compute_sdf_Hg(o, fx["hg.R_last"], fx["hg.t_last"], delta.q.rotation_matrix, delta.t, fx["hg.obs"], k, robust_kernel="huber")
    assert abs(oE - hE) > 1e-3 * abs(hE)                          # the two kernels differ on this data
    assert hg_errors(H, g, E, H2, g2, E2)[0] <= 1.0 and close(E, E2, 1e-5)
    assert hg_errors(H, g, E, oH, og, oE)[0] <= 1.0 and close(E, oE, 1e-4)
    with pytest.raises(NotImplementedError):
        SDFTracker(m, argparse.Namespace(sdf=dict(robust_kernel="cauchy", robust_k=k, subsample=0.5), rgb=None, iter_config=[])) \
            .compute_sdf_Hg(0, last, delta, obs)
