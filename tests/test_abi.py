"""CPU: the C-ABI library loads and exports every symbol include/difusion_b200.h declares; host-side logic."""
import re
from pathlib import Path

import numpy as np
import pytest
import torch

from conftest import GOLDEN, ROOT


def _header_functions():
    txt = (ROOT / "include" / "difusion_b200.h").read_text()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(dif_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    from difusion_b200 import _lib, build
    build.build()
    L = _lib.lib()
    names = _header_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(L, n), f"{n} declared in the header but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature"
    assert set(_lib.SIGNATURES) == set(names)
    assert L.dif_abi_version() == _lib.ABI_VERSION


def test_workspace_queries_need_no_gpu():
    from difusion_b200 import _lib
    L = _lib.lib()
    assert L.dif_decoder_prepared_bytes() >= 4 * 49408 * 2
    assert L.dif_encoder_prepared_bytes() >= 4 * 26048
    assert L.dif_integrate_scratch_bytes(30000) > 30000 * 12
    assert L.dif_integrate_persist_bytes(1920000, 1 << 16) > 1920000 * 4
    assert L.dif_mesh_decode_scratch_bytes(1000, 4) >= 1000 * 512 * 4
    assert L.dif_icp_scratch_bytes(30000) > 0


def test_struct_layout_matches_header():
    from difusion_b200 import _lib
    import ctypes
    # 6 pointers + int64 + 3 int32 + 3 float + float + int32 + 2 float + 2 int32 + pointer = 48 + 8 + 12 + 12 + 4 + 4 + 8 + 8 + 8 = 112
    assert ctypes.sizeof(_lib.MapView) == 152 and _lib.MapView.scalar_division_mode.offset == 144 and _lib.MapView.xchg_slots.offset == 104 and _lib.MapView.latent_stride.offset == 112
    assert _lib.MapView.shard_block_log2.offset == 116 and _lib.MapView.row_of_slot.offset == 120 and _lib.MapView.row_capacity.offset == 136
    assert _lib.MapView.capacity.offset == 48 and _lib.MapView.nx.offset == 56 and _lib.MapView.bound_min.offset == 68


def test_struct_layout_against_the_c_compiler(tmp_path):
    """sizeof / offsetof of every struct of the header as gcc lays them out == the ctypes mirrors in _lib.py."""
    import ctypes
    import subprocess
    from difusion_b200 import _lib
    probes = {"dif_map_view": (_lib.MapView, ["capacity", "bound_min", "xchg_slots", "row_capacity", "scalar_division_mode"]),
              "dif_frame_params": (_lib.FrameParams, ["pose"]),
              "dif_gn_level": (_lib.GnLevel, ["cur_grad", "h"]),
              "dif_gn_group": (_lib.GnGroup, ["kind", "level"]),
              "dif_gn_problem": (_lib.GnProblem, ["n_obs", "huber_k", "level", "intr", "K", "Kinv", "min_grad_scale", "rgb_weight", "n_groups",
                                                  "group", "last_pose", "init_delta"]),
              "dif_gn_result": (_lib.GnResult, ["energy", "last_iter", "n_rgb"])}
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{ROOT / "include" / "difusion_b200.h"}"', "int main(void) {"]
    for name, (_, fields) in probes.items():
        lines.append(f'printf("{name} %zu\\n", sizeof({name}));')
        for f in fields:
            lines.append(f'printf("{name}.{f} %zu\\n", offsetof({name}, {f}));')
    lines.append("return 0; }")
    src = tmp_path / "probe.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "probe"
    subprocess.check_call(["gcc", "-std=c11", str(src), "-o", str(exe)])
    got = dict(ln.split() for ln in subprocess.check_output([str(exe)], text=True).splitlines())
    for name, (cls, fields) in probes.items():
        assert int(got[name]) == ctypes.sizeof(cls), (name, got[name], ctypes.sizeof(cls))
        for f in fields:
            assert int(got[f"{name}.{f}"]) == getattr(cls, f).offset, (name, f)


def test_no_cpu_fallback():
    from difusion_b200 import _lib
    from difusion_b200.network import utility as net_util
    from difusion_b200.system.map import DenseIndexedMap
    from difusion_b200 import synthetic as S
    model, _ = net_util.load_model(str(GOLDEN / "weights.npz"))
    with pytest.raises(_lib.DifusionLibraryError):
        DenseIndexedMap(model, S.scene_S0().map_args(), 29, torch.device("cpu"))


def test_fold_matches_oracle(oracle_weights):
    from difusion_b200 import weights
    dec, enc = weights.load_npz_state(GOLDEN / "weights.npz")
    blob = weights.fold_decoder(dec)
    o = oracle_weights.dec
    ref = np.concatenate([o.W[0].numpy().ravel(), o.b[0].numpy(), o.W[1].numpy().ravel(), o.b[1].numpy(), o.W[2].numpy().ravel(),
                          o.b[2].numpy(), o.W[3].numpy().ravel(), o.b[3].numpy(), o.W[4].numpy().ravel(), o.b[4].numpy(),
                          o.Wu.numpy().ravel(), o.bu.numpy()])
    assert blob.shape == ref.shape and np.abs(blob - ref).max() < 1e-6
    eblob = weights.fold_encoder(enc)
    e = oracle_weights.enc
    eref = np.concatenate([np.concatenate([e.W[k].numpy().ravel(), e.b[k].numpy()]) for k in range(4)])
    assert eblob.shape == eref.shape and np.abs(eblob - eref).max() < 1e-6


def test_isometry_matches_oracle_se3():
    from difusion_b200.utils.motion_util import Isometry
    from oracle import dif_oracle as O
    xi = np.array([0.02, -0.01, 0.03, 0.05, -0.04, 0.02])
    iso = Isometry.from_twist(xi)
    R, t = O.se3_exp(xi)
    assert np.allclose(iso.q.rotation_matrix, R, atol=1e-12) and np.allclose(iso.t, t, atol=1e-12)
    a = Isometry.from_twist(xi * 0.3)
    ab = iso.dot(a)
    assert np.allclose(ab.matrix, iso.matrix @ a.matrix, atol=1e-12)
    assert np.allclose(iso.inv().dot(iso).matrix, np.eye(4), atol=1e-12)
    p = np.random.default_rng(0).normal(size=(5, 3))
    assert np.allclose(iso @ p, p @ R.T + t)
    assert np.allclose((iso @ torch.from_numpy(p).float()).numpy(), (p @ R.T + t).astype(np.float32), atol=1e-6)


def test_gauss_newton_update_step_matches_numpy():
    """The fp64 algebra of the device-driven Gauss-Newton loop (csrc/gn.cu: partial-pivot 6x6 solve, exp map with the left Jacobian,
    composition), evaluated on the host through dif_debug_gn_step, against numpy.linalg.solve + Isometry.from_twist @ delta
    (reference system/tracker.py:270-272, utils/motion_util.py:205-229,277-278)."""
    import ctypes
    from difusion_b200 import _lib
    from difusion_b200.utils.motion_util import Isometry
    L = _lib.lib()
    rng = np.random.default_rng(5)
    for trial in range(20):
        J = rng.normal(size=(200, 6)) * rng.uniform(0.1, 30.0, size=6)
        H = J.T @ J / 200
        g = J.T @ rng.normal(size=200) / 200 * (1e-9 if trial == 0 else 1.0)          # trial 0: |phi| ~ 1e-10, the small-angle branch
        d0 = Isometry.from_twist(rng.normal(size=6) * 0.05)
        delta = np.concatenate([d0.q.rotation_matrix.ravel(), d0.t]).astype(np.float64)
        buf = (ctypes.c_double * 12)(*delta)
        rc = L.dif_debug_gn_step((ctypes.c_double * 36)(*H.ravel()), (ctypes.c_double * 6)(*g), buf)
        assert rc == 0
        ref = Isometry.from_twist(np.linalg.solve(H, -g)) @ d0
        got = np.asarray(buf[:])
        assert np.abs(got[:9].reshape(3, 3) - ref.q.rotation_matrix).max() < 1e-12 and np.abs(got[9:] - ref.t).max() < 1e-12
    sing = np.zeros((6, 6)); sing[0, 0] = 1.0
    assert L.dif_debug_gn_step((ctypes.c_double * 36)(*sing.ravel()), (ctypes.c_double * 6)(*np.ones(6)), buf) == _lib.GN_SINGULAR
