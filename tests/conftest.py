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
