"""CPU: oracle/imgproc_oracle.c (gradient_xy, rgb_odometry, unproject_depth) against golden vectors produced by EXECUTING the
unmodified reference CUDA extension on a B200 (tests/golden/ref_ext_photo.npz, ref_ext_unproject.npz; generator
tests/golden/make_golden_gpu.py) - bit-exact - plus the numpy restatement of compute_rgb_Hg on top of it."""
import numpy as np

from conftest import GOLDEN
from oracle import imgproc_oracle as O


def _bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def test_gradient_xy_bit_exact():
    fx = np.load(GOLDEN / "ref_ext_photo.npz")
    g = O.gradient_xy(fx["cur_i"])
    assert np.isnan(g[0]).all() and np.isnan(g[:, -1]).all() and not np.isnan(g[1:-1, 1:-1]).any()
    assert np.array_equal(_bits(g), _bits(fx["grad"]))


def test_rgb_odometry_bit_exact():
    fx = np.load(GOLDEN / "ref_ext_photo.npz")
    f, J = O.rgb_odometry(fx["prev_i"], fx["prev_d"], fx["cur_i"], fx["cur_d"], fx["grad"], fx["intr"].tolist(), fx["krkinv"].tolist(),
                          fx["kt"].tolist(), float(fx["min_grad_scale"]), float(fx["max_depth_delta"]))
    v = ~np.isnan(fx["f"])
    assert v.sum() > 10000 and np.array_equal(v, ~np.isnan(f))
    assert np.array_equal(_bits(f[v]), _bits(fx["f"][v])) and np.array_equal(_bits(J[v]), _bits(fx["J"][v]))
    f2, = O.rgb_odometry(fx["prev_i"], fx["prev_d"], fx["cur_i"], fx["cur_d"], fx["grad"], fx["intr"].tolist(), fx["krkinv"].tolist(),
                         fx["kt"].tolist(), float(fx["min_grad_scale"]), float(fx["max_depth_delta"]), compute_J=False)
    assert np.array_equal(_bits(f2), _bits(f))


def test_unproject_bit_exact():
    fx = np.load(GOLDEN / "ref_ext_unproject.npz")
    pc = O.unproject_depth(fx["depth"], *[float(v) for v in fx["intr"]])
    assert np.array_equal(_bits(pc), _bits(fx["pc"]))


def test_compute_rgb_Hg_properties():
    """The normal equations built from the pinned residuals/Jacobians: symmetric PSD H, and one Gauss-Newton step lowers the energy."""
    fx = np.load(GOLDEN / "ref_ext_photo.npz")
    a = (fx["prev_i"], fx["prev_d"], fx["cur_i"], fx["cur_d"], fx["grad"], fx["intr"], fx["K"])
    H, g, E, M = O.compute_rgb_Hg(*a, fx["Rd"], fx["td"], 1e-5, 0.2, 500.0)
    assert M > 10000 and np.allclose(H, H.T) and np.linalg.eigvalsh(H).min() > 0
    for kind, k in (("huber", 0.004), ("tukey", 0.02)):
        Hr, gr, Er, Mr = O.compute_rgb_Hg(*a, fx["Rd"], fx["td"], 1e-5, 0.2, 500.0, kind, k)
        assert Mr == M and Er < E and np.linalg.eigvalsh(Hr).min() >= -1e-9
    _, _, E0, _ = O.compute_rgb_Hg(*a, fx["Rd"], fx["td"], 1e-5, 0.2, 500.0, no_grad=True)
    assert E0 == E
    # the fixture's delta pose was perturbed by (4, -3, 2) mm; the unperturbed pose must have a lower photometric energy
    _, _, E_true, _ = O.compute_rgb_Hg(*a, fx["Rd"], fx["td"] - np.array([0.004, -0.003, 0.002]), 1e-5, 0.2, 500.0, no_grad=True)
    assert E_true < 0.7 * E
