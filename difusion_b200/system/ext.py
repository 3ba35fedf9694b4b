"""Drop-in replacements for the reference's native op table ``system.ext`` (ext/__init__.py:15-44) on the hot path:
``marching_cubes_interp`` (mc.cpp:3-16), ``groupby_sum`` (indexing.cpp:4) and, from the image side,
``unproject_depth`` (imgproc.cu:26-44), ``gradient_xy`` and ``rgb_odometry`` (photometric.cu:80-138).  Same argument meaning and
return values; torch tensors in, torch tensors out; the work is done by libdifusion_b200.so on the current stream.
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


def gradient_xy(cur_intensity: torch.Tensor):
    """(H,W) f32 -> (H,W,2) f32 Sobel gradients / 8, NaN on the border (photometric.cu:3-22,80-93)."""
    _check_input(cur_intensity, "cur_intensity")
    h, w = cur_intensity.shape
    out = torch.empty((h, w, 2), dtype=torch.float32, device=cur_intensity.device)
    _lib.check(_lib.lib().dif_gradient_xy(cur_intensity.data_ptr(), h, w, out.data_ptr(), _lib.stream_ptr(cur_intensity.device)), "dif_gradient_xy")
    return out


def rgb_odometry(prev_intensity, prev_depth, cur_intensity, cur_depth, cur_dIdxy, intr, krkinv_data, kt_data,
                 min_grad_scale: float, max_depth_delta: float, compute_J: bool):
    """-> [f_img (H,W)] or [f_img, J_img (H,W,6)]; NaN in f_img marks rejected pixels (photometric.cu:24-78,95-138)."""
    for t, nm in ((prev_intensity, "prev_intensity"), (prev_depth, "prev_depth"), (cur_intensity, "cur_intensity"),
                  (cur_depth, "cur_depth"), (cur_dIdxy, "cur_dIdxy")):
        _check_input(t, nm)
    h, w = cur_intensity.shape
    dev = cur_intensity.device
    f_img = torch.empty((h, w), dtype=torch.float32, device=dev)
    J_img = torch.empty((h, w, 6), dtype=torch.float32, device=dev) if compute_J else None
    _lib.check(_lib.lib().dif_rgb_odometry(prev_intensity.data_ptr(), prev_depth.data_ptr(), cur_intensity.data_ptr(), cur_depth.data_ptr(),
                                           cur_dIdxy.data_ptr(), h, w, _lib.host_floats(intr), _lib.host_floats(krkinv_data),
                                           _lib.host_floats(kt_data), float(min_grad_scale), float(max_depth_delta), f_img.data_ptr(),
                                           _lib.ptr(J_img), _lib.stream_ptr(dev)), "dif_rgb_odometry")
    return [f_img, J_img] if compute_J else [f_img]


def unproject_depth(depth: torch.Tensor, fx: float, fy: float, cx: float, cy: float):
    """(H,W) f32 depth (NaN = invalid) -> (H,W,3) camera-frame points (imgproc.cu:5-44); invalid pixels are NaN."""
    _check_input(depth, "depth")
    h, w = depth.shape
    pc = torch.empty((h, w, 3), dtype=torch.float32, device=depth.device)
    _lib.check(_lib.lib().dif_unproject_depth(depth.data_ptr(), h, w, float(fx), float(fy), float(cx), float(cy), pc.data_ptr(),
                                              _lib.stream_ptr(depth.device)), "dif_unproject_depth")
    return pc


_BOX_MAX_CELLS = 1 << 27                 # 2 cm cells: a ~10 m cube of bounding box (16 MB bitmap + 16 MB ranks)
_box_scratch = {}                        # (device, stream) -> [zero-filled scratch, points it was sized for, cell budget]


def point_box_filter(points: torch.Tensor, normals: torch.Tensor, voxel_size: float, deferred: bool = False):
    """tracker.point_box_filter (tracker.py:13-23): (N,3),(N,3) -> (M,3),(M,3) per-cell means, cells in ascending key order.
    Rows with a NaN coordinate are skipped.  deferred=True: no host sync - returns (out_p (N,3), out_n (N,3), n_out device int32[1]);
    the first n_out rows are valid, n_out < 0 flags a bounding box beyond the scratch's cell budget."""
    _check_input(points, "points")
    _check_input(normals, "normals")
    L, dev, n = _lib.lib(), points.device, points.size(0)
    key = (dev, torch.cuda.current_stream(dev).cuda_stream)
    out_p = torch.empty((n, 3), dtype=torch.float32, device=dev)
    out_n = torch.empty((n, 3), dtype=torch.float32, device=dev)
    n_out = torch.zeros(1, dtype=torch.int32, device=dev)
    cells = _BOX_MAX_CELLS
    for attempt in range(2):
        sc = _box_scratch.get(key)
        if sc is None or sc[1] < n or sc[2] < cells:
            cap = max(n, 1 << 17, sc[1] if sc else 0)
            sc = [torch.zeros(L.dif_box_filter_scratch_bytes(cap, max(cells, sc[2] if sc else 0)), dtype=torch.uint8, device=dev), cap,
                  max(cells, sc[2] if sc else 0)]
            _box_scratch[key] = sc
        _lib.check(L.dif_point_box_filter(points.data_ptr(), normals.data_ptr(), n, float(voxel_size), sc[2], out_p.data_ptr(),
                                          out_n.data_ptr(), n_out.data_ptr(), sc[0].data_ptr(), sc[0].numel(), _lib.stream_ptr(dev)), "dif_point_box_filter")
        if deferred:
            return out_p, out_n, n_out
        m = int(n_out.item())                    # host sync: output shape (the reference syncs at tracker.py:18 and inside unique)
        if m >= 0:
            return out_p[:m], out_n[:m]
        # the frame's bounding box does not fit the bitmap: size it from the actual extent (tracker.py:18 adds 16 cells per axis) and retry once
        ext_ = (points.amax(dim=0) - points.amin(dim=0)).double().cpu().numpy()
        need = 1
        for e in ext_:
            need *= int(e / float(voxel_size)) + 18
        cells = 1 << max(need - 1, 1).bit_length()
        if attempt == 1 or cells > (1 << 29):
            break
    raise RuntimeError("point_box_filter: the frame's bounding box exceeds the cell budget of the filter scratch")


_KNN_MAX_CELLS = 1 << 23                 # cell edge = search radius: 5 cm cells cover a ~10 m cube of bounding box
_KNN_CELLS_LIMIT = 1 << 30               # hard ceiling of the on-demand growth below (a ~50 m cube at 5 cm)
_knn_scratch = {}                        # (device, stream) -> [zero-filled scratch, points it was sized for, cell budget]


def _knn_key(dev):
    return (dev, torch.cuda.current_stream(dev).cuda_stream)


def _knn_buffers(dev, n, min_cells=_KNN_MAX_CELLS):
    """Scratch of the neighbour grid for `n` points: per (device, stream), grown on demand.  -> (tensor, cell budget)"""
    L = _lib.lib()
    key = _knn_key(dev)
    sc = _knn_scratch.get(key)
    if sc is None or sc[1] < n or sc[2] < min_cells:
        cap = max(n, 1 << 17, sc[1] if sc else 0)
        cells = max(min_cells, sc[2] if sc else 0)
        sc = [torch.zeros(L.dif_knn_scratch_bytes(cap, cells), dtype=torch.uint8, device=dev), cap, cells]
        _knn_scratch[key] = sc
    return sc[0], sc[2]


def _knn_cells_needed(input_pc: torch.Tensor, cell: float) -> int:
    """Cells of edge `cell` covering the finite points' bounding box (with the kernels' one-cell margin): the overflow path only."""
    p = input_pc[:, :3]
    ok = torch.isfinite(p).all(dim=1)
    if not bool(ok.any()):
        return 1
    p = p[ok]
    ext_ = (p.amax(dim=0) - p.amin(dim=0)).double().cpu().numpy()
    n = 1
    for e in ext_:
        n *= int(e / cell) + 4
    return n


def _knn_call(fn_name, input_pc, cell, call):
    """Run a neighbour-grid kernel; if the cloud's bounding box does not fit the cell budget of the scratch (a single far or noisy depth
    pixel is enough: the reference's kd-tree has no such limit), size the budget from the actual extent and run again."""
    dev, n = input_pc.device, input_pc.size(0)
    sc, cells = _knn_buffers(dev, n)
    status = call(sc, cells)
    if int(status.item()):
        need = _knn_cells_needed(input_pc, cell)
        if need > _KNN_CELLS_LIMIT:
            raise RuntimeError(f"{fn_name}: the cloud's bounding box needs {need} neighbour cells (limit {_KNN_CELLS_LIMIT})")
        sc, cells = _knn_buffers(dev, n, 1 << max(need - 1, 1).bit_length())
        status = call(sc, cells)
        if int(status.item()):
            raise RuntimeError(f"{fn_name}: the cloud's bounding box exceeds the neighbour grid budget")


def remove_radius_outlier(input_pc: torch.Tensor, nb_points: int, radius: float, status_out: list = None):
    """(N,4) [or (N,3)] f32 -> (N,) bool: the nb_points-th nearest point (self included) lies within `radius` (pcproc.cu:172-196).
    Rows with a NaN coordinate are "no point": they are nobody's neighbour and their mask is False (an un-compacted cloud gives the
    same masks for its valid rows as the compacted one).  status_out (list): no host sync - the overflow flag (device int32) is
    appended for the caller to check later instead of growing the cell budget here."""
    _check_input(input_pc, "input_pc")
    dev, n = input_pc.device, input_pc.size(0)
    mask = torch.empty(n, dtype=torch.uint8, device=dev)

    def call(sc, cells):
        status = torch.zeros(1, dtype=torch.int32, device=dev)
        _lib.check(_lib.lib().dif_remove_radius_outlier(input_pc.data_ptr(), input_pc.size(1), n, int(nb_points), float(radius), cells,
                                                        mask.data_ptr(), status.data_ptr(), sc.data_ptr(), sc.numel(), _lib.stream_ptr(dev)),
                   "dif_remove_radius_outlier")
        return status
    if status_out is not None:
        status_out.append(call(*_knn_buffers(dev, n)))
    else:
        _knn_call("remove_radius_outlier", input_pc, float(radius), call)
    return mask.view(torch.bool)


def estimate_normals(input_pc: torch.Tensor, max_nn: int, radius: float, cam_xyz, status_out: list = None):
    """(N,4) [or (N,3)] f32 -> (N,3) f32 PCA normals over the <= max_nn-1 nearest neighbours within `radius`, oriented towards
    cam_xyz; NaN rows where fewer than 5 neighbours exist (pcproc.cu:107-170,198-220).  NaN input rows are "no point" (NaN normal);
    status_out as in remove_radius_outlier."""
    _check_input(input_pc, "input_pc")
    dev, n = input_pc.device, input_pc.size(0)
    normals = torch.empty((n, 3), dtype=torch.float32, device=dev)

    def call(sc, cells):
        status = torch.zeros(1, dtype=torch.int32, device=dev)
        _lib.check(_lib.lib().dif_estimate_normals(input_pc.data_ptr(), input_pc.size(1), n, int(max_nn), float(radius), _lib.host_floats(cam_xyz),
                                                   cells, normals.data_ptr(), status.data_ptr(), sc.data_ptr(), sc.numel(),
                                                   _lib.stream_ptr(dev)), "dif_estimate_normals")
        return status
    if status_out is not None:
        status_out.append(call(*_knn_buffers(dev, n)))
    else:
        _knn_call("estimate_normals", input_pc, 0.5 * float(radius), call)             # two-ring search: the grid's cells are radius / 2
    return normals
