"""Golden fixtures for the latent-optimisation branch (reference system/map.py:453-516, 80-117), produced by running the
UNMODIFIED reference (oracle/ref_shim.py) on CPU tensors.  Build-container only.

    python tests/golden/make_golden_opt.py

    latent_opt.npz    OptimizeProcess.do_optimize (map.py:80-117) on seeded rows / samples: inputs + optimised rows, with and without
                      code regularisation; oracle/dif_oracle.optimize_latent_rows is asserted against it in the same run
    s0_optimize.npz   scene S0, 3 frames through the reference DenseIndexedMap.integrate_keyframe(do_optimize=True); the sample noise
                      of map.py:486 (torch.randn on the device) is replaced for that call by a reproducible stream - the k-th request of
                      n samples returns numpy default_rng(1000 + k).standard_normal(n) - so that the product can be fed the same
                      noise: latents, voxel_optimized, updated ids
"""
import argparse
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from difusion_b200 import synthetic as S          # noqa: E402
from oracle import dif_oracle as O               # noqa: E402
from oracle import ref_shim                       # noqa: E402

OUT = Path(__file__).resolve().parent
torch.manual_seed(0)
torch.set_num_threads(8)

N_ITERS, LAMBDA = 5, 1.0e-2


def main():
    ref = ref_shim.load_reference()
    model, _ = ref_shim.load_reference_model()
    W = O.load_weights_npz(OUT / "weights.npz")

    # ---------------------------------------------------------------- row-level: the reference's own do_optimize
    g = torch.Generator().manual_seed(77)
    U, n = 300, 6000
    lat = torch.randn(U, 29, generator=g) * 0.3
    inv = torch.randint(0, U, (n,), generator=g)
    sdf = torch.randn(n, generator=g) * 0.05
    sdf[::11] *= 6.0                                          # some targets beyond the +-0.2 clamp
    rel = torch.rand(n, 3, generator=g) - 0.5 + sdf.unsqueeze(-1) * torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=1)
    fx = dict(latent=lat.numpy(), inv=inv.numpy(), sdf=sdf.numpy(), rel=rel.numpy(), n_iters=np.int64(N_ITERS), code_reg_lambda=np.float64(LAMBDA))
    for reg in (False, True):
        args = argparse.Namespace(optim_n_iters=N_ITERS, code_regularization=reg, code_reg_lambda=LAMBDA)
        out = ref.map.OptimizeProcess.do_optimize(model.decoder, args, lat.clone(), inv.clone(), sdf.clone(), rel.clone()).detach().numpy()
        mine = O.optimize_latent_rows(W.dec, lat.numpy(), inv.numpy(), sdf.numpy(), rel.numpy(), N_ITERS, reg, LAMBDA)
        d = np.abs(out - mine).max()
        print(f"do_optimize reg={reg}: moved rows by max {np.abs(out - lat.numpy()).max():.4f}; oracle vs reference max abs {d:.2e}")
        assert d <= 2e-5, d
        fx["out_reg" if reg else "out"] = out
    np.savez(OUT / "latent_opt.npz", **fx)

    # ---------------------------------------------------------------- map-level: integrate_keyframe(do_optimize=True), noise off
    sc = S.scene_S0()
    args = sc.map_args()
    args.encoder_count_th = 40.0                              # PLIVoxes reach the optimisation threshold within the 3 frames
    args.optim_n_iters, args.code_regularization, args.code_reg_lambda = N_ITERS, True, LAMBDA
    rmap = ref.map.DenseIndexedMap(model, argparse.Namespace(**vars(args)), 29, torch.device("cpu"))
    real_randn = torch.randn
    noise = S.ReproducibleNoise()
    fm = dict(encoder_count_th=np.float64(args.encoder_count_th), n_iters=np.int64(N_ITERS), code_reg_lambda=np.float64(LAMBDA))
    captured = {}
    real_do = ref.map.OptimizeProcess.do_optimize

    def spy_do_optimize(decoder, args, latent_vecs_unique, latent_id_inv_mapping, gathered_sdf, gathered_relative_xyz):
        captured.update(lat0=latent_vecs_unique.detach().clone().numpy(), inv=latent_id_inv_mapping.clone().numpy(),
                        sdf=gathered_sdf.clone().numpy(), rel=gathered_relative_xyz.clone().numpy(),
                        ids=rmap.optimize_result_set.latent_ids.clone().numpy())
        return real_do(decoder, args, latent_vecs_unique, latent_id_inv_mapping, gathered_sdf, gathered_relative_xyz)
    rmap.optimize_process.do_optimize = spy_do_optimize
    for f, yaw in enumerate((0.0, 0.06, 0.12)):
        R, t = S.yaw_pose(yaw)
        pc, nc = S.frame_points(sc, R, t)
        xw, nw = S.to_world(pc, nc, R, t)
        torch.randn = noise.torch_randn
        try:
            rmap.integrate_keyframe(torch.from_numpy(xw), torch.from_numpy(nw), do_optimize=True, async_optimize=False)
        finally:
            torch.randn = real_randn
        nocc = rmap.n_occupied
        if f == 0:          # what the reference handed to do_optimize (map.py:495-507): the gather of step 3, to be matched exactly
            fm.update({"f0.gather.ids": captured["ids"].astype(np.int32), "f0.gather.inv": captured["inv"].astype(np.int32),
                       "f0.gather.sdf": captured["sdf"], "f0.gather.rel": captured["rel"], "f0.gather.lat0": captured["lat0"]})
        fm[f"f{f}.yaw"] = np.float64(yaw)
        fm[f"f{f}.n_occupied"] = np.int64(nocc)
        fm[f"f{f}.latent"] = rmap.latent_vecs.numpy()[:nocc].copy()
        fm[f"f{f}.obs_count"] = rmap.voxel_obs_count.numpy()[:nocc].copy()
        fm[f"f{f}.optimized"] = rmap.voxel_optimized.numpy()[:nocc].copy()
        fm[f"f{f}.updated_vec_id"] = np.sort(rmap.mesh_cache.updated_vec_id.numpy()).astype(np.int32)
        print(f"[s0_optimize] frame {f}: n_occ={nocc} optimised so far={int(rmap.voxel_optimized[:nocc].sum())} "
              f"obs>=th={(rmap.voxel_obs_count[:nocc] >= args.encoder_count_th).sum().item()}")
    assert int(rmap.voxel_optimized[:nocc].sum()) > 50
    np.savez_compressed(OUT / "s0_optimize.npz", **fm)


if __name__ == "__main__":
    main()
