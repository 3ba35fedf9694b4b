"""ctypes binding of libdifusion_b200.so (C ABI: include/difusion_b200.h).

The product path has NO CPU fallback: if the CUDA library is missing or fails to load, every entry point raises.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_HERE = Path(__file__).resolve().parent
LIB_PATH = Path(os.environ["DIF_LIB_PATH"]) if os.environ.get("DIF_LIB_PATH") else _HERE / "libdifusion_b200.so"   # override: kernel A/B builds (tools/)

ABI_VERSION = 3
DIF_STAT_COUNT = 12
STAT_N_KEPT, STAT_N_NEW, STAT_N_SAMPLES, STAT_N_UPDATED, STAT_N_OCCUPIED, STAT_FLAGS, STAT_N_FOCUSED, STAT_N_XCHG, STAT_SEQ, STAT_N_ROWS = range(10)
FRAME_HEADER_FLOATS, FRAME_POINT_FLOATS = 32, 9           # DIF_FRAME_HEADER_FLOATS / DIF_FRAME_POINT_FLOATS
FRAME_TRACK, FRAME_INTEGRATE = 1, 2
LATENT_ROW_FLOATS = 32       # the map stores 128-byte latent rows (dif_map_view.latent_stride); columns 29..31 are padding
FRAME_RESULT_BYTES = 44 * 8 + DIF_STAT_COUNT * 4


class MapView(C.Structure):
    """struct dif_map_view"""
    _fields_ = [("indexer", C.c_void_p), ("latent_vecs", C.c_void_p), ("latent_vecs_pos", C.c_void_p),
                ("voxel_obs_count", C.c_void_p), ("slot_dirty", C.c_void_p), ("n_occupied", C.c_void_p),
                ("capacity", C.c_int64), ("nx", C.c_int32), ("ny", C.c_int32), ("nz", C.c_int32),
                ("bound_min", C.c_float * 3), ("voxel_size", C.c_float), ("prune_min_vox_obs", C.c_int32),
                ("ignore_count_th", C.c_float), ("encoder_count_th", C.c_float),
                ("shard_rank", C.c_int32), ("shard_world", C.c_int32), ("xchg_slots", C.c_void_p), ("latent_stride", C.c_int32),
                ("shard_block_log2", C.c_int32), ("row_of_slot", C.c_void_p), ("n_rows", C.c_void_p), ("row_capacity", C.c_int64),
                ("scalar_division_mode", C.c_int32), ("reserved_", C.c_int32)]


class FrameParams(C.Structure):
    """struct dif_frame_params (the per-frame block the kernels read from DEVICE memory)"""
    _fields_ = [("n_points", C.c_int32), ("seq", C.c_int32), ("reserved", C.c_int32 * 2), ("pose", C.c_float * 24)]


GN_MAX_TERMS, GN_MAX_GROUPS, GN_MAX_LEVELS = 4, 8, 4
GN_TERM_SDF, GN_TERM_RGB = 0, 1
GN_CONTINUE, GN_BREAK, GN_EMPTY, GN_SINGULAR = 1, 2, 3, 4


class GnLevel(C.Structure):
    """struct dif_gn_level"""
    _fields_ = [("prev_i", C.c_void_p), ("prev_d", C.c_void_p), ("cur_i", C.c_void_p), ("cur_d", C.c_void_p), ("cur_grad", C.c_void_p),
                ("h", C.c_int32), ("w", C.c_int32)]


class GnGroup(C.Structure):
    """struct dif_gn_group"""
    _fields_ = [("n_iters", C.c_int32), ("n_terms", C.c_int32), ("kind", C.c_int32 * GN_MAX_TERMS), ("level", C.c_int32 * GN_MAX_TERMS)]


class GnProblem(C.Structure):
    """struct dif_gn_problem"""
    _fields_ = [("obs_xyz", C.c_void_p), ("n_obs", C.c_int64), ("huber_k", C.c_float), ("n_levels", C.c_int32),
                ("level", GnLevel * GN_MAX_LEVELS), ("intr", C.c_float * 4), ("K", C.c_double * 9), ("Kinv", C.c_double * 9),
                ("min_grad_scale", C.c_float), ("max_depth_delta", C.c_float), ("rgb_robust", C.c_int32), ("rgb_robust_k", C.c_float),
                ("rgb_weight", C.c_float), ("n_groups", C.c_int32), ("group", GnGroup * GN_MAX_GROUPS),
                ("last_pose", C.c_double * 12), ("init_delta", C.c_double * 12)]


class GnResult(C.Structure):
    """struct dif_gn_result"""
    _fields_ = [("delta", C.c_double * 12), ("energy", C.c_double), ("last_iter", C.c_int32), ("status", C.c_int32),
                ("empty_term", C.c_int32), ("n_iterations", C.c_int32), ("n_sdf", C.c_int32), ("n_rgb", C.c_int32)]


_P, _I64, _I32, _F, _SZ = C.c_void_p, C.c_int64, C.c_int32, C.c_float, C.c_size_t
_MV = C.POINTER(MapView)
_FP = C.POINTER(C.c_float)          # small HOST float arrays (intrinsics, K R K^-1, K t, bound_min)

# name -> (restype, argtypes); this table is also what tests/test_abi.py checks against include/difusion_b200.h
SIGNATURES = {
    "dif_abi_version": (C.c_int, []),
    "dif_shard_owner": (C.c_int, [_I64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "dif_shard_xchg_bytes": (_SZ, [_I64, C.c_int]),
    "dif_shard_pack": (C.c_int, [_MV, _P, _I64, _P, _P]),
    "dif_shard_unpack": (C.c_int, [_MV, _P, _I64, _P, _P]),
    "dif_shard_select_points": (C.c_int, [_MV, _P, _I64, _P, _P, _P, _P]),
    "dif_profile_hook": (C.c_int, [C.c_int, _P, _P]),
    "dif_launch_count": (C.c_uint64, [C.c_int]),
    "dif_debug_tc_timing": (C.c_int, [_P]),
    "dif_last_error": (C.c_char_p, []),
    "dif_decoder_prepared_bytes": (_SZ, []),
    "dif_encoder_prepared_bytes": (_SZ, []),
    "dif_prepare_decoder": (C.c_int, [_P, _P, _P]),
    "dif_prepare_encoder": (C.c_int, [_P, _P, _P]),
    "dif_integrate_persist_bytes": (_SZ, [_I64, _I64]),
    "dif_integrate_scratch_bytes": (_SZ, [_I64]),
    "dif_integrate": (C.c_int, [_MV, _P, _P, _P, _I64, _P, _P, _P, _SZ, _P, _SZ, _P, _P]),
    "dif_decode": (C.c_int, [_P, _P, C.c_int, _P, _P, _I64, _P, _F, _P, _P, _P, _P, _P]),
    "dif_encode": (C.c_int, [_P, _P, _I64, _P, _P]),
    "dif_map_query": (C.c_int, [_MV, _P, _I64, _P, _P, _P, _P]),
    "dif_icp_scratch_bytes": (_SZ, [_I64]),
    "dif_icp_linearize": (C.c_int, [_MV, _P, _P, _I64, _P, _P, _F, C.c_int, _P, _SZ, _P, _P]),
    "dif_frame": (C.c_int, [_MV, _P, _P, _P, _I64, _P, _F, C.c_int, _P, _P, _SZ, _P, _SZ, _P, _SZ, _P, _P]),
    "dif_mesh_select_scratch_bytes": (_SZ, [_I64, _I64]),
    "dif_mesh_select": (C.c_int, [_MV, _P, _I64, _P, _P, _P, _P, _P, _SZ, _P]),
    "dif_mesh_decode_scratch_bytes": (_SZ, [_I64, C.c_int]),
    "dif_mesh_decode": (C.c_int, [_MV, _P, _P, _I64, C.c_int, C.c_int, _P, _P, _P, _SZ, _P, _P]),
    "dif_marching_cubes": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, _P, _I64, _P, _I64, _P, _P, C.c_int, _F, _P, _P, _P, _I64, _P, _P]),
    "dif_groupby_sum": (C.c_int, [_P, _P, _I64, _I32, _I64, _P, _P, _P]),
    "dif_unproject_depth": (C.c_int, [_P, C.c_int, C.c_int, _F, _F, _F, _F, _P, _P]),
    "dif_box_filter_scratch_bytes": (_SZ, [_I64, _I64]),
    "dif_point_box_filter": (C.c_int, [_P, _P, _I64, _F, _I64, _P, _P, _P, _P, _SZ, _P]),
    "dif_knn_scratch_bytes": (_SZ, [_I64, _I64]),
    "dif_remove_radius_outlier": (C.c_int, [_P, C.c_int, _I64, C.c_int, _F, _I64, _P, _P, _P, _SZ, _P]),
    "dif_estimate_normals": (C.c_int, [_P, C.c_int, _I64, C.c_int, _F, _FP, _I64, _P, _P, _P, _SZ, _P]),
    "dif_gradient_xy": (C.c_int, [_P, C.c_int, C.c_int, _P, _P]),
    "dif_rgb_odometry": (C.c_int, [_P, _P, _P, _P, _P, C.c_int, C.c_int, _FP, _FP, _FP, _F, _F, _P, _P, _P]),
    "dif_rgb_scratch_bytes": (_SZ, []),
    "dif_rgb_linearize": (C.c_int, [_P, _P, _P, _P, _P, C.c_int, C.c_int, _FP, _FP, _FP, _F, _F, C.c_int, _F, _F, C.c_int, _P, _SZ, _P, _P]),
    "dif_latent_grad": (C.c_int, [_P, _P, _P, _P, _P, _I64, _I64, _P, _P, _P]),
    "dif_debug_gn_step": (C.c_int, [C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "dif_gn_scratch_bytes": (_SZ, [_I64]),
    "dif_gauss_newton": (C.c_int, [_MV, _P, C.POINTER(GnProblem), _P, _SZ, _P, C.POINTER(GnResult), _P]),
    "dif_mesh_cache_scratch_bytes": (_SZ, [_I64, _I64]),
    "dif_mesh_cache_merge": (C.c_int, [_P, _P, _P, _I64, _P, _P, _P, _I64, _F, C.POINTER(C.c_float), _I64, _P, _P, _P, _P, _P, _SZ, _P]),
}

_lib = None


class DifusionLibraryError(RuntimeError):
    pass


def lib():
    """The loaded library; raises (loudly) when it has not been built - there is no fallback path."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise DifusionLibraryError(
                f"{LIB_PATH} not found: build it with `python -m difusion_b200.build` (nvcc, sm_100a). "
                "difusion_b200 has no CPU or PyTorch fallback.")
        try:
            h = C.CDLL(str(LIB_PATH))
        except OSError as e:
            raise DifusionLibraryError(f"cannot load {LIB_PATH}: {e}") from e
        for name, (res, args) in SIGNATURES.items():
            f = getattr(h, name)
            f.restype, f.argtypes = res, args
        if h.dif_abi_version() != ABI_VERSION:
            raise DifusionLibraryError("ABI version mismatch")
        _lib = h
    return _lib


def check(rc: int, what: str):
    if rc != 0:
        msg = lib().dif_last_error().decode() if rc == -3 else {-1: "invalid argument", -2: "workspace too small"}.get(rc, "?")
        raise DifusionLibraryError(f"{what} failed with code {rc}: {msg}")


def ptr(t):
    """Device pointer of a (contiguous) torch tensor, or None."""
    if t is None:
        return None
    assert t.is_contiguous(), "difusion_b200 kernels need contiguous tensors"
    return t.data_ptr()


def host_floats(values):
    """ctypes float array for the small host-side parameter vectors of the ABI."""
    vals = [float(v) for v in values]
    return (C.c_float * len(vals))(*vals)


def stream_ptr(device=None):
    import torch
    return torch.cuda.current_stream(device).cuda_stream
