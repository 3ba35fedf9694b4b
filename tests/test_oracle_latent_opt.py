"""CPU: the latent-optimisation restatement (oracle/dif_oracle.optimize_latent_rows, reference system/map.py:80-117) against the
fixture produced by the UNMODIFIED reference's OptimizeProcess.do_optimize (tests/golden/make_golden_opt.py)."""
import numpy as np
import pytest

from conftest import GOLDEN


@pytest.mark.parametrize("reg", [False, True])
def test_optimize_latent_rows_matches_reference(oracle_weights, reg):
    from oracle import dif_oracle as O
    fx = np.load(GOLDEN / "latent_opt.npz")
    out = O.optimize_latent_rows(oracle_weights.dec, fx["latent"], fx["inv"], fx["sdf"], fx["rel"], int(fx["n_iters"]), reg, float(fx["code_reg_lambda"]))
    ref = fx["out_reg" if reg else "out"]
    assert np.abs(ref - fx["latent"]).max() > 0.01                       # the optimiser moved the rows
    assert np.abs(out - ref).max() <= 2e-5


def test_chunked_loss_counts_the_regulariser_per_chunk(oracle_weights):
    """forward_model calls loss_func once per chunk (utility.py:86-118), so the regulariser's gradient is added n_chunks times."""
    from oracle import dif_oracle as O
    fx = np.load(GOLDEN / "latent_opt.npz")
    one = O.optimize_latent_rows(oracle_weights.dec, fx["latent"], fx["inv"][:600], fx["sdf"][:600], fx["rel"][:600], 2, True, 5.0)
    two = O.optimize_latent_rows(oracle_weights.dec, fx["latent"], fx["inv"][:600], fx["sdf"][:600], fx["rel"][:600], 2, True, 5.0, max_sample=300)
    assert np.abs(one - two).max() > 1e-4
