"""GPU: latent optimisation (reference system/map.py:453-516, 80-117) through the mirror + dif_latent_grad, against
fixtures produced by the UNMODIFIED reference (tests/golden/make_golden_opt.py).

Each piece is pinned separately, because Adam's update lr * m / (sqrt(v) + eps) is scale-free: an element whose gradient changes sign
moves by up to +-lr whatever the gradient's size, so a 1e-5 difference of the STARTING rows (our tensor-core encoder against the
reference's fp32 one, inside the 1e-4 bar) can move a few elements by 1e-3 after 5 steps.  Therefore:
  * the optimiser on identical inputs (do_optimize's own arguments)          -> stated bar |a-b| <= 1e-4 + 1e-4|b| (measured 1.3e-6);
  * one backward pass (dif_latent_grad) against torch autograd               -> 1e-4 of the gradient's scale;
  * the gather of step 3 (which samples, which rows, inverse map, targets)   -> exact / 1e-6 against what the reference passed on;
  * the whole integrate_keyframe(do_optimize=True) over 3 frames: integer state exact, latents at the stated bar on >= 97 % of the
    elements and within three of the five Adam steps (3e-2) on all (measured: 98.6 %, 1.0e-2 = one step on a single element)."""
import numpy as np
import pytest
import torch

from conftest import GOLDEN, frac_off

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


def _map(dev, args):
    from difusion_b200.network import utility as net_util
    from difusion_b200.system.map import DenseIndexedMap
    model, _ = net_util.load_model(str(GOLDEN / "weights.npz"))
    return DenseIndexedMap(model, args, 29, dev)


@pytest.mark.parametrize("reg", [False, True])
def test_optimize_latent_rows_against_reference(dev, reg):
    from difusion_b200 import synthetic as S
    fx = np.load(GOLDEN / "latent_opt.npz")
    args = S.scene_S0().map_args()
    args.optim_n_iters, args.code_regularization, args.code_reg_lambda = int(fx["n_iters"]), reg, float(fx["code_reg_lambda"])
    m = _map(dev, args)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    out = m.optimize_latent_rows(t(fx["latent"]), t(fx["inv"]), t(fx["sdf"]), t(fx["rel"])).cpu().numpy()
    ref = fx["out_reg" if reg else "out"]
    d = np.abs(out - ref)
    print(f"[latent opt] reg={reg}: max |a-b| {d.max():.2e}, off the 1e-4 bar: {frac_off(out, ref):.2e}")
    assert frac_off(out, ref) <= 2e-3 and d.max() <= 2e-3


def test_latent_grad_matches_autograd(dev):
    """dif_latent_grad (one backward pass) against torch autograd over the oracle's decoder on the same samples."""
    from difusion_b200 import _lib, synthetic as S
    from oracle import dif_oracle as O
    fx = np.load(GOLDEN / "latent_opt.npz")
    W = O.load_weights_npz(GOLDEN / "weights.npz")
    m = _map(dev, S.scene_S0().map_args())
    lat = torch.from_numpy(fx["latent"]).requires_grad_(True)
    inv, sdf, rel = torch.from_numpy(fx["inv"]), torch.from_numpy(fx["sdf"]), torch.from_numpy(fx["rel"])
    p_sdf, p_std = O.decoder_forward(W.dec, lat[inv], rel)
    ll = -torch.distributions.Normal(loc=torch.clamp(p_sdf, -0.2, 0.2), scale=p_std).log_prob(torch.clamp(sdf, -0.2, 0.2))
    loss = ll.sum() / inv.shape[0]
    loss.backward()
    g = torch.zeros(lat.shape, dtype=torch.float32, device=dev)
    lo = torch.zeros(1, dtype=torch.float64, device=dev)
    d_lat, d_inv, d_rel, d_sdf = (x.detach().to(dev).contiguous() for x in (lat, inv, rel, sdf))
    _lib.check(_lib.lib().dif_latent_grad(m._prep.decoder.data_ptr(), d_lat.data_ptr(), d_inv.data_ptr(), d_rel.data_ptr(), d_sdf.data_ptr(),
                                          inv.shape[0], inv.shape[0], g.data_ptr(), lo.data_ptr(), _lib.stream_ptr(dev)), "dif_latent_grad")
    ref = lat.grad.numpy()
    err = np.abs(g.cpu().numpy() - ref).max()
    assert err <= 1e-4 * np.abs(ref).max() + 1e-9, (err, np.abs(ref).max())
    assert abs(float(lo.item()) - float(loss)) <= 1e-5 * abs(float(loss))


def test_integrate_keyframe_with_do_optimize_against_reference(dev):
    from difusion_b200 import synthetic as S
    fx = np.load(GOLDEN / "s0_optimize.npz")
    sc = S.scene_S0()
    args = sc.map_args()
    args.encoder_count_th = float(fx["encoder_count_th"])
    args.optim_n_iters, args.code_regularization, args.code_reg_lambda = int(fx["n_iters"]), True, float(fx["code_reg_lambda"])
    m = _map(dev, args)
    m.optim_noise_fn = S.ReproducibleNoise()                             # the stream the fixture's reference run was fed (map.py:486)
    seen = {}
    inner = m.optimize_latent_rows

    def spy(lat, inv, sdf, rel):
        if not seen:
            seen.update(lat0=lat.cpu().numpy(), inv=inv.cpu().numpy(), sdf=sdf.cpu().numpy(), rel=rel.cpu().numpy())
        return inner(lat, inv, sdf, rel)
    m.optimize_latent_rows = spy
    for f in range(3):
        R, t = S.yaw_pose(float(fx[f"f{f}.yaw"]))
        pc, nc = S.frame_points(sc, R, t)
        xw, nw = S.to_world(pc, nc, R, t)
        m.integrate_keyframe(torch.from_numpy(xw).to(dev), torch.from_numpy(nw).to(dev), do_optimize=True)
        if f == 0:      # the gather of step 3 == what the reference handed to do_optimize
            assert np.array_equal(seen["inv"], fx["f0.gather.inv"]) and np.array_equal(seen["sdf"], fx["f0.gather.sdf"])
            assert np.abs(seen["rel"] - fx["f0.gather.rel"]).max() <= 1e-6
            assert np.array_equal(np.nonzero(fx["f0.optimized"])[0], fx["f0.gather.ids"])
            d0 = np.abs(seen["lat0"] - fx["f0.gather.lat0"])
            print(f"[s0 optimize] starting rows (encoder output): max |a-b| {d0.max():.2e}")
            assert np.all(d0 <= 1e-4 + 1e-4 * np.abs(fx["f0.gather.lat0"]))
        n = int(fx[f"f{f}.n_occupied"])
        assert m.n_occupied == n
        assert np.array_equal(m.voxel_optimized.cpu().numpy()[:n], fx[f"f{f}.optimized"])
        assert np.array_equal(m.voxel_obs_count.cpu().numpy()[:n], fx[f"f{f}.obs_count"])
        assert np.array_equal(np.sort(m.mesh_cache.updated_vec_id.cpu().numpy()), fx[f"f{f}.updated_vec_id"])
        lat, ref = m.latent_vecs.cpu().numpy()[:n], fx[f"f{f}.latent"]
        d = np.abs(lat - ref)
        print(f"[s0 optimize] frame {f}: {int(fx[f'f{f}.optimized'].sum())} optimised PLIVoxes, max |latent - ref| {d.max():.2e}, off the bar {frac_off(lat, ref):.2e}")
        assert frac_off(lat, ref) <= 3e-2 and d.max() <= 3e-2
    with pytest.raises(NotImplementedError):
        m.integrate_keyframe(torch.from_numpy(xw).to(dev), torch.from_numpy(nw).to(dev), do_optimize=True, async_optimize=True)
