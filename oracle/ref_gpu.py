"""Run the UNMODIFIED reference (huangjh-pub/di-fusion, pytorch/) through its own torch-CUDA path on the GPU box.

TEST / BASELINE INFRASTRUCTURE ONLY.  Used by tests/test_ref_gpu_path.py (live GPU-side parity of integrate_keyframe and
compute_sdf_Hg) and by bench.py's `reference_gpu` extra (SURVEY 8(d) "Timing the reference", item 2: the reference's GPU path
timed beside the headline step).  Nothing under difusion_b200/ may import it.

What runs is the reference's own Python (system/map.py:340-519 integrate_keyframe, system/tracker.py:174-218 compute_sdf_Hg,
network/*.py) staged byte for byte by oracle/build_ref.stage_python() under oracle/_ref/pytorch/ (git-ignored; in the build
container /root/reference/pytorch is used directly), with its native op table `system.ext` (ext/__init__.py:15-44, a JIT
`cpp_extension.load` of the CUDA sources) bound to the same sources compiled ahead of time by oracle/build_ref.build()
(oracle/_ref/<name>/<name>.so).  Substitutions, all for packages that are absent from this image, none on the timed path's math:

* ``open3d``                      -> empty stub module                 (map.py:6, GUI / mesh container only)
* ``pyquaternion.Quaternion``     -> oracle.ref_shim.Quaternion        (motion_util.py:2)
* ``torch_scatter.scatter_mean``  -> index_add_ mean on the same device (tracker.py:14,21-22; not on the timed path)
* ``np.product``                  -> ``np.prod``                       (map.py:178,201,407; removed in NumPy 2)
* ``mp.set_start_method``         -> no-op                             (map.py:172 forces 'forkserver' on the whole process)
"""
from __future__ import annotations

import sys
import time
import types
from pathlib import Path

import numpy as np
import torch

from . import build_ref
from .ref_shim import Quaternion

_LOADED = None


def root() -> Path:
    return build_ref.REF_PY if build_ref.REF_PY.exists() else build_ref.PY_OUT


def available() -> bool:
    """The staged reference Python + checkpoint and the four compiled extensions are present."""
    have_py = build_ref.REF_PY.exists() or build_ref.python_available()
    return have_py and all(build_ref.available(n) for n in ("marching_cubes", "indexing", "imgproc", "pcproc"))


def _scatter_mean(src, index, dim=0):
    n = int(index.max()) + 1
    out = torch.zeros((n, src.size(1)), dtype=src.dtype, device=src.device).index_add_(0, index, src)
    cnt = torch.zeros((n,), dtype=src.dtype, device=src.device).index_add_(0, index, torch.ones_like(index, dtype=src.dtype))
    return out / cnt.unsqueeze(-1)


def load():
    """-> namespace(map, tracker, net_util, motion_util, di_decoder, di_encoder) of the reference's own modules, `system.ext` bound
    to the reference's own compiled CUDA extensions."""
    global _LOADED
    if _LOADED is not None:
        return _LOADED
    if not available():
        raise RuntimeError("reference GPU path unavailable: run __graft_entry__.build() where /root/reference is mounted")
    sys.dont_write_bytecode = True
    sys.path.insert(0, str(root()))
    sys.modules.setdefault("open3d", types.ModuleType("open3d"))
    pq = types.ModuleType("pyquaternion"); pq.Quaternion = Quaternion
    sys.modules.setdefault("pyquaternion", pq)
    ts = types.ModuleType("torch_scatter"); ts.scatter_mean = _scatter_mean
    sys.modules.setdefault("torch_scatter", ts)
    if not hasattr(np, "product"):
        np.product = np.prod
    import torch.multiprocessing as mp
    mp.set_start_method = lambda *a, **k: None

    mc, ix = build_ref.load_module("marching_cubes"), build_ref.load_module("indexing")
    im, pp = build_ref.load_module("imgproc"), build_ref.load_module("pcproc")
    ext = types.ModuleType("system.ext")                   # ext/__init__.py:15-44, same names
    ext.marching_cubes_interp = mc.marching_cubes_sparse_interp
    ext.unproject_depth, ext.rgb_odometry, ext.gradient_xy = im.unproject_depth, im.rgb_odometry, im.gradient_xy
    ext.compute_normal_weight, ext.compute_normal_weight_robust, ext.filter_depth = \
        im.compute_normal_weight, im.compute_normal_weight_robust, im.filter_depth
    ext.pack_batch, ext.groupby_sum = ix.pack_batch, ix.groupby_sum
    ext.remove_radius_outlier, ext.estimate_normals = pp.remove_radius_outlier, pp.estimate_normals
    import system                                           # the reference's namespace package
    sys.modules["system.ext"] = ext
    system.ext = ext

    import json
    import network.di_decoder as di_decoder
    import network.di_encoder as di_encoder
    import network.utility as net_util
    from system import map as refmap
    from system import tracker as reftracker
    from utils import motion_util
    _LOADED = types.SimpleNamespace(map=refmap, tracker=reftracker, net_util=net_util, di_decoder=di_decoder, di_encoder=di_encoder,
                                    motion_util=motion_util, json=json)
    return _LOADED


def load_model(device):
    """The shipped checkpoint (ckpt/default) as the reference's ``Networks`` object on `device` (network/utility.py:23-63)."""
    ref = load()
    ck = root() / "ckpt" / "default"
    hyper = ref.json.load(open(ck / "hyper.json"))
    model = ref.net_util.Networks()
    model.decoder = ref.di_decoder.Model(hyper["code_length"], **hyper["network_specs"]).to(device)
    model.encoder = ref.di_encoder.Model(**hyper["encoder_specs"]).to(device)
    model.decoder.load_state_dict(torch.load(ck / "model_300.pth.tar", map_location=device)["model_state"])
    model.encoder.load_state_dict(torch.load(ck / "encoder_300.pth.tar", map_location=device)["model_state"])
    model.eval()
    return model


TRACK_ARGS = dict(sdf=dict(robust_kernel="huber", robust_k=5.0, subsample=0.5),
                  rgb=dict(weight=500.0, robust_kernel=None, robust_k=0.01, min_grad_scale=0.0, max_depth_delta=0.2),
                  iter_config=[{"n": 50, "type": [["sdf"]]}])


class RefStream:
    """The headline step (ICP linearisation of the frame against the map + integrate_keyframe) through the reference's own classes."""

    def __init__(self, map_args, device):
        import argparse
        self.ref = load()
        self.dev = device
        self.model = load_model(device)
        self.map = self.ref.map.DenseIndexedMap(self.model, argparse.Namespace(**vars(map_args)), 29, device)
        self.tracker = self.ref.tracker.SDFTracker(self.map, argparse.Namespace(**TRACK_ARGS))
        self.Iso = self.ref.motion_util.Isometry

    def pose(self, R, t):
        return self.Iso.from_matrix(np.block([[np.asarray(R, float), np.asarray(t, float).reshape(3, 1)], [np.zeros((1, 3)), np.ones((1, 1))]]))

    def step(self, f, pc_cam, xw, nw, R, t):
        """pc_cam / xw / nw: device tensors.  Returns (H, g, E) of the linearisation (None for frame 0)."""
        out = None
        if f >= 1:
            out = self.tracker.compute_sdf_Hg(0, self.pose(R, t), self.Iso(), pc_cam, no_grad=False)
        self.map.integrate_keyframe(xw, nw)
        return out


def time_stream(map_args, frames, device, n_steps: int, warmup: int = 1):
    """Seconds for `n_steps` headline steps on a fresh reference map (after `warmup` untimed frames on a scratch map: cuDNN/cuBLAS
    handles, allocator), inputs device-resident, wall clock with a device sync on both sides.  Returns (seconds, RefStream)."""
    def dev_frames(n):
        return [(torch.from_numpy(fr["pc"]).to(device), torch.from_numpy(fr["xw"]).to(device), torch.from_numpy(fr["nw"]).to(device),
                 fr["R"], fr["t"]) for fr in frames[:n]]
    d = dev_frames(max(n_steps, warmup))
    scratch = RefStream(map_args, device)
    for f in range(min(warmup, len(d))):
        scratch.step(f, *d[f])
    del scratch
    rs = RefStream(map_args, device)
    torch.cuda.synchronize(device)
    t0 = time.perf_counter()
    for f in range(n_steps):
        rs.step(f, *d[f])
    torch.cuda.synchronize(device)
    return time.perf_counter() - t0, rs
