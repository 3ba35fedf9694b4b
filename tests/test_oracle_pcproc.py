"""CPU: oracle/pcproc_oracle.py (remove_radius_outlier, estimate_normals restated over scipy's exact k-NN) against golden vectors
produced by EXECUTING the unmodified reference kd-tree extension on a B200 (tests/golden/ref_ext_pcproc.npz)."""
import numpy as np
import pytest

from conftest import GOLDEN
from oracle import pcproc_oracle as P


@pytest.mark.parametrize("tag", ["r5", "r8"])
def test_knn_ops_against_executed_reference(tag):
    fx = np.load(GOLDEN / "ref_ext_pcproc.npz")
    cloud, radius = fx["cloud"], float(fx[f"{tag}.radius"])
    mask = P.remove_radius_outlier(cloud, 16, radius)
    assert (mask != fx[f"{tag}.mask"]).mean() <= 5e-4            # a float64 tree vs fp32 distances: only 1-ulp radius ties may differ
    ref_n = fx[f"{tag}.normals"]
    n = P.estimate_normals(cloud[fx[f"{tag}.mask"]], 16, 2 * radius, [0.0, 0.0, 0.0])
    na, nb = np.isnan(n[:, 0]), np.isnan(ref_n[:, 0])
    assert (na != nb).mean() <= 1e-3
    both = ~na & ~nb
    cosang = np.abs((n[both].astype(np.float64) * ref_n[both]).sum(1))
    assert (cosang < np.cos(1e-3)).mean() <= 5e-3               # closed-form fp32 eigenvector vs LAPACK, equal-distance ties
    assert ((ref_n[both] * cloud[fx[f"{tag}.mask"]][both, :3]).sum(1) <= 0).all()


def test_edge_cases():
    one = np.zeros((1, 4), np.float32)
    assert not P.remove_radius_outlier(one, 16, 0.05)[0] and np.isnan(P.estimate_normals(one, 16, 0.1, [0, 0, 0])).all()
    rng = np.random.default_rng(0)
    plane = np.concatenate([rng.uniform(-0.05, 0.05, (200, 2)), np.full((200, 1), 2.0), np.zeros((200, 1))], 1).astype(np.float32)
    assert P.remove_radius_outlier(plane, 16, 0.05).all()
    n = P.estimate_normals(plane, 16, 0.1, [0, 0, 0])
    assert np.allclose(n, [0, 0, -1], atol=1e-5)
