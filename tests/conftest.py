import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))
GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def close(a, b, tol=1e-4):
    """The parity criterion of BASELINE.json north_star / SURVEY 8(d): |a-b| <= tol + tol*|b|."""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return bool(np.all(np.abs(a - b) <= tol + tol * np.abs(b)))


def frac_off(a, b, tol=1e-4):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float((np.abs(a - b) > tol + tol * np.abs(b)).mean())


def hg_errors(H, g, E, H_ref, g_ref, E_ref, tol=1e-4):
    """Normal equations of compute_sdf_Hg (tracker.py:209-216) against a reference evaluation.  H = mean(w J J^T) is a Gram
    matrix and g = mean(J w r): element (a, b) is a sum of ~30 k products whose magnitudes are bounded by sqrt(H_aa H_bb)
    (Cauchy-Schwarz), resp. sqrt(H_aa E), not by |H_ab| - an off-diagonal element can be 3000x smaller than that scale through
    cancellation (fixture s1_map: H_03 = 176 vs sqrt(H_00 H_33) = 81 000).  Measured on that fixture: the CPU restatement of the
    reference (oracle.compute_sdf_Hg, torch fp32) against the EXECUTED reference (torch fp32) sits at 3.5 x 'tol + tol |b|' on
    such an element and at 0.13 x the Gram-scaled bound (tests/test_oracle_golden.py asserts both numbers' order of magnitude):
    the verbatim bar is below what two fp32 evaluations of the reference's own code agree to.
    Returned: (gram, strict) = max over elements of |a - b| / (tol + tol * scale) with scale = the Gram bound (the bar the tests
    assert, identical to the stated element-wise bar on the diagonal) and scale = |b| (the stated bar verbatim, recorded)."""
    H, g, H_ref, g_ref = (np.asarray(x, np.float64) for x in (H, g, H_ref, g_ref))
    d = np.sqrt(np.abs(np.diag(H_ref)))
    gram = max(float((np.abs(H - H_ref) / (tol + tol * np.outer(d, d))).max()),
               float((np.abs(g - g_ref) / (tol + tol * d * np.sqrt(abs(float(E_ref))))).max()))
    strict = max(float((np.abs(H - H_ref) / (tol + tol * np.abs(H_ref))).max()), float((np.abs(g - g_ref) / (tol + tol * np.abs(g_ref))).max()))
    return gram, strict


@pytest.fixture(scope="session")
def golden():
    class G:
        def __getitem__(self, name):
            return np.load(GOLDEN / f"{name}.npz")
    return G()


@pytest.fixture(scope="session")
def oracle_weights():
    from oracle import dif_oracle as O
    return O.load_weights_npz(GOLDEN / "weights.npz")


def fixture_args(fx):
    import argparse
    return argparse.Namespace(bound_min=fx["bound_min"].tolist(), bound_max=fx["bound_max"].tolist(),
                              voxel_size=float(fx["voxel_size"]), prune_min_vox_obs=int(fx["prune_min_vox_obs"]),
                              ignore_count_th=float(fx["ignore_count_th"]), encoder_count_th=float(fx["encoder_count_th"]),
                              optim_n_iters=0)
