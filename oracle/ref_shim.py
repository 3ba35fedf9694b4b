"""Import the UNMODIFIED reference (huangjh-pub/di-fusion, /root/reference/pytorch) on CPU tensors.

TEST INFRASTRUCTURE ONLY.  This module exists to (a) pin the restatement in
``oracle/dif_oracle.py`` against the real reference code and (b) generate the
golden fixtures under ``tests/golden/`` (see ``tests/golden/make_golden.py``).
It only works in the build container, where ``/root/reference`` is mounted; the
GPU box never imports it.  Nothing under ``difusion_b200/`` may import it.

The reference has no CPU mode (``torch.cuda.Stream()`` in the map ctor,
``.cuda()`` in the tracker, CUDA-only extensions).  The shim below is the
smallest set of substitutions that lets its own Python run on CPU tensors:

* ``open3d``                       -> empty stub module        (map.py:6, GUI only)
* ``system.ext``                   -> CPU stand-ins            (ext/__init__.py:15-44)
    - ``groupby_sum``              -> ``index_add_``           (indexing.cu:59-109)
    - ``marching_cubes_interp``    -> ``oracle.mc_oracle``     (mc_interp_kernel.cu:7-382)
* ``torch.cuda.Stream/stream/synchronize`` -> no-ops           (map.py:232,625-626)
* ``np.product``                   -> ``np.prod``              (map.py:178,201,407; NumPy>=2)
* ``pyquaternion.Quaternion``      -> minimal shim             (motion_util.py:2)
* ``torch_scatter.scatter_mean``   -> ``index_add_`` mean      (tracker.py:14,21-22)
* ``Tensor.cuda()``                -> identity                 (tracker.py:197)
"""
from __future__ import annotations

import contextlib
import os
import sys
import types
from pathlib import Path

import numpy as np
import torch

REFERENCE_ROOT = Path(os.environ.get("DIF_REFERENCE_ROOT", "/root/reference/pytorch"))


def reference_available() -> bool:
    return (REFERENCE_ROOT / "system" / "map.py").exists()


# --------------------------------------------------------------------------- pyquaternion shim
class Quaternion:
    """Minimal stand-in for pyquaternion.Quaternion (w, x, y, z), enough for utils/motion_util.py."""

    def __init__(self, *args, **kwargs):
        if "matrix" in kwargs:
            m = np.asarray(kwargs["matrix"], dtype=float)
            self.q = self._from_matrix(m[:3, :3])
        elif "axis" in kwargs:
            axis = np.asarray(kwargs["axis"], dtype=float)
            axis = axis / np.linalg.norm(axis)
            ang = kwargs.get("radians", None)
            if ang is None:
                ang = np.deg2rad(kwargs.get("degrees", 0.0))
            self.q = np.concatenate([[np.cos(ang / 2)], np.sin(ang / 2) * axis])
        elif "array" in kwargs:
            self.q = np.asarray(kwargs["array"], dtype=float).copy()
        elif len(args) == 4:
            self.q = np.asarray(args, dtype=float)
        elif len(args) == 1 and isinstance(args[0], Quaternion):
            self.q = args[0].q.copy()
        elif len(args) == 1:
            self.q = np.asarray(args[0], dtype=float).copy()
        else:
            self.q = np.array([1.0, 0.0, 0.0, 0.0])

    @staticmethod
    def _from_matrix(R):
        t = np.trace(R)
        if t > 0:
            s = np.sqrt(t + 1.0) * 2
            q = [0.25 * s, (R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s]
        elif R[0, 0] > R[1, 1] and R[0, 0] > R[2, 2]:
            s = np.sqrt(1.0 + R[0, 0] - R[1, 1] - R[2, 2]) * 2
            q = [(R[2, 1] - R[1, 2]) / s, 0.25 * s, (R[0, 1] + R[1, 0]) / s, (R[0, 2] + R[2, 0]) / s]
        elif R[1, 1] > R[2, 2]:
            s = np.sqrt(1.0 + R[1, 1] - R[0, 0] - R[2, 2]) * 2
            q = [(R[0, 2] - R[2, 0]) / s, (R[0, 1] + R[1, 0]) / s, 0.25 * s, (R[1, 2] + R[2, 1]) / s]
        else:
            s = np.sqrt(1.0 + R[2, 2] - R[0, 0] - R[1, 1]) * 2
            q = [(R[1, 0] - R[0, 1]) / s, (R[0, 2] + R[2, 0]) / s, (R[1, 2] + R[2, 1]) / s, 0.25 * s]
        q = np.asarray(q, dtype=float)
        return q / np.linalg.norm(q)

    @property
    def rotation_matrix(self):
        w, x, y, z = self.q / np.linalg.norm(self.q)
        return np.array([
            [1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
            [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
            [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])

    @property
    def transformation_matrix(self):
        m = np.eye(4)
        m[:3, :3] = self.rotation_matrix
        return m

    @property
    def inverse(self):
        w, x, y, z = self.q
        return Quaternion(np.array([w, -x, -y, -z]) / np.dot(self.q, self.q))

    def rotate(self, v):
        return self.rotation_matrix @ np.asarray(v, dtype=float)

    def __mul__(self, o):
        a, b = self.q, o.q
        return Quaternion(np.array([
            a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3],
            a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2],
            a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1],
            a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0]]))

    def __repr__(self):
        return f"Quaternion{tuple(self.q)}"


def _scatter_mean(src, index, dim=0):
    n = int(index.max()) + 1
    out = torch.zeros((n, src.size(1)), dtype=src.dtype).index_add_(0, index, src)
    cnt = torch.zeros((n,), dtype=src.dtype).index_add_(0, index, torch.ones_like(index, dtype=src.dtype))
    return out / cnt.unsqueeze(-1)


def _groupby_sum(values, indices, C):
    """CPU stand-in for ext/indexing/indexing.cu:59-109 (only the sum is consumed by utility.py:200-206)."""
    C = int(C)
    s = torch.zeros(C, values.size(1), dtype=torch.float32).index_add_(0, indices, values)
    c = torch.zeros(C, dtype=torch.int32).index_add_(0, indices, torch.ones_like(indices, dtype=torch.int32))
    return s, c * values.size(1)      # the kernel bumps the count once per column (indexing.cu:70)


def _marching_cubes_interp(indexer, valid_blocks, vec_batch_mapping, cube_sdf, cube_std,
                           max_n_triangles, n_xyz, max_std):
    from oracle import mc_oracle
    tri, fid, std = mc_oracle.marching_cubes_interp(
        indexer.numpy(), valid_blocks.numpy(), vec_batch_mapping.numpy(),
        cube_sdf.numpy(), cube_std.numpy(), int(max_n_triangles), list(n_xyz), float(max_std))
    return torch.from_numpy(tri), torch.from_numpy(fid), torch.from_numpy(std)


_LOADED = None


def load_reference():
    """Returns a namespace with the reference's modules (map, tracker, utility, decoder, encoder, motion_util)."""
    global _LOADED
    if _LOADED is not None:
        return _LOADED
    if not reference_available():
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT}")
    sys.dont_write_bytecode = True                      # the reference tree is read-only
    sys.path.insert(0, str(REFERENCE_ROOT))

    sys.modules.setdefault("open3d", types.ModuleType("open3d"))
    pq = types.ModuleType("pyquaternion"); pq.Quaternion = Quaternion
    sys.modules.setdefault("pyquaternion", pq)
    ts = types.ModuleType("torch_scatter"); ts.scatter_mean = _scatter_mean
    sys.modules.setdefault("torch_scatter", ts)

    ext = types.ModuleType("system.ext")
    ext.groupby_sum = _groupby_sum
    ext.marching_cubes_interp = _marching_cubes_interp
    for name in ("unproject_depth", "remove_radius_outlier", "estimate_normals", "rgb_odometry", "gradient_xy"):
        setattr(ext, name, None)
    import system  # noqa: the reference's namespace package
    sys.modules["system.ext"] = ext
    system.ext = ext

    torch.cuda.Stream = lambda *a, **k: None
    torch.cuda.stream = lambda s: contextlib.nullcontext()
    torch.cuda.synchronize = lambda *a, **k: None
    torch.Tensor.cuda = lambda self, *a, **k: self
    if not hasattr(np, "product"):
        np.product = np.prod

    import json
    import network.di_decoder as di_decoder
    import network.di_encoder as di_encoder
    import network.utility as net_util
    from system import map as refmap
    from system import tracker as reftracker
    from utils import motion_util

    ns = types.SimpleNamespace(map=refmap, tracker=reftracker, net_util=net_util, di_decoder=di_decoder,
                               di_encoder=di_encoder, motion_util=motion_util, json=json)
    _LOADED = ns
    return ns


def load_reference_model():
    """The shipped checkpoint (ckpt/default) as a reference ``Networks`` object on CPU."""
    ref = load_reference()
    ck = REFERENCE_ROOT / "ckpt" / "default"
    hyper = ref.json.load(open(ck / "hyper.json"))
    model = ref.net_util.Networks()
    model.decoder = ref.di_decoder.Model(hyper["code_length"], **hyper["network_specs"])
    model.encoder = ref.di_encoder.Model(**hyper["encoder_specs"])
    model.decoder.load_state_dict(torch.load(ck / "model_300.pth.tar", map_location="cpu")["model_state"])
    model.encoder.load_state_dict(torch.load(ck / "encoder_300.pth.tar", map_location="cpu")["model_state"])
    model.eval()
    return model, hyper
