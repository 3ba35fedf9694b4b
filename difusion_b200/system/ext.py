"""Drop-in replacements for the reference's native op table ``system.ext`` (ext/__init__.py:15-44) on the hot path:
``marching_cubes_interp`` (mc.cpp:3-16) and ``groupby_sum`` (indexing.cpp:4).  Same argument meaning and return
values; torch tensors in, torch tensors out; the work is done by libdifusion_b200.so on the current stream.
"""
from __future__ import annotations

import sys

import torch

from .. import _lib


def _check_input(t: torch.Tensor, name: str):
    # mirrors CHECK_INPUT (mc_data.cuh:7-9): RuntimeError for non-CUDA / non-contiguous inputs
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor")
    if not t.is_contiguous():
        raise RuntimeError(f"{name} must be contiguous")


def marching_cubes_interp(indexer: torch.Tensor, valid_blocks: torch.Tensor, vec_batch_mapping: torch.Tensor,
                          cube_sdf: torch.Tensor, cube_std: torch.Tensor, max_n_triangles: int, n_xyz, max_std: float):
    """-> [triangles (T,3,3) f32 voxel units, triangle_flatten_id (T,) i64, triangle_std (T,3) f32]"""
    for t, nm in ((indexer, "indexer"), (valid_blocks, "valid_blocks"), (cube_sdf, "cube_sdf"), (cube_std, "cube_std"),
                  (vec_batch_mapping, "vec_batch_mapping")):
        _check_input(t, nm)
    assert max_n_triangles > 0
    assert indexer.dtype == torch.int64 and valid_blocks.dtype == torch.int64 and vec_batch_mapping.dtype == torch.int32
    dev = cube_sdf.device
    r = cube_sdf.size(1) // 2
    tri = torch.empty((max_n_triangles, 3, 3), dtype=torch.float32, device=dev)
    fid = torch.empty((max_n_triangles,), dtype=torch.int64, device=dev)
    std = torch.empty((max_n_triangles, 3), dtype=torch.float32, device=dev)
    count = torch.zeros(1, dtype=torch.int32, device=dev)
    _lib.check(_lib.lib().dif_marching_cubes(
        indexer.data_ptr(), int(n_xyz[0]), int(n_xyz[1]), int(n_xyz[2]), valid_blocks.data_ptr(), valid_blocks.size(0),
        vec_batch_mapping.data_ptr(), vec_batch_mapping.size(0), cube_sdf.data_ptr(), cube_std.data_ptr(), r, float(max_std),
        tri.data_ptr(), fid.data_ptr(), std.data_ptr(), int(max_n_triangles), count.data_ptr(), _lib.stream_ptr(dev)), "dif_marching_cubes")
    n = int(count.item())                      # the reference syncs here too (mc_interp_kernel.cu:367-369)
    if n < max_n_triangles:
        return [tri[:n], fid[:n], std[:n]]
    sys.stderr.write(f"Warning from marching cube: the max triangle number is too small {n} vs {max_n_triangles}\n")
    return [tri, fid, std]


def groupby_sum(values: torch.Tensor, indices: torch.Tensor, C: int):
    """-> [sum (C,L) f32, count (C,) i32]; the count is bumped once per column like the reference kernel (indexing.cu:70)."""
    _check_input(values, "values")
    _check_input(indices, "indices")
    C = int(C)
    n, Lc = values.size(0), values.size(1)
    s = torch.zeros((C, Lc), dtype=torch.float32, device=values.device)
    c = torch.zeros((C,), dtype=torch.int32, device=values.device)
    _lib.check(_lib.lib().dif_groupby_sum(values.data_ptr(), indices.data_ptr(), n, Lc, C, s.data_ptr(), c.data_ptr(),
                                          _lib.stream_ptr(values.device)), "dif_groupby_sum")
    return [s, c]
