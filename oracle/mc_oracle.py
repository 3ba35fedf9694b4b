"""ctypes front-end of oracle/mc_oracle.c (ORACLE - test infrastructure only, see that file's header)."""
from __future__ import annotations

import ctypes
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB = None


def build(force: bool = False) -> Path:
    out = _HERE / "_build" / "libdif_oracle.so"
    srcs = [_HERE / "mc_oracle.c", _HERE / "imgproc_oracle.c", _HERE / "Makefile"]
    if force or not out.exists() or out.stat().st_mtime < max(s.stat().st_mtime for s in srcs):
        subprocess.check_call(["make", "-C", str(_HERE), "-s"] + (["-B"] if force else []))
    return out


def _lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(str(build()))
        f = _LIB.dif_oracle_marching_cubes
        f.restype = ctypes.c_int64
        P = ctypes.c_void_p
        f.argtypes = [P, P, ctypes.c_int64, P, ctypes.c_int64, P, P, ctypes.c_int,
                      ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_int64, P, P, P]
    return _LIB


def marching_cubes_interp(indexer, valid_blocks, vec_batch_mapping, cube_sdf, cube_std,
                          max_n_triangles: int, n_xyz, max_std: float):
    """Same contract as reference system.ext.marching_cubes_interp (mc.cpp:3-16), numpy in / numpy out.
    Returns (tri (T,3,3) f32 in voxel units, flatten_id (T,) i64, tri_std (T,3) f32)."""
    indexer = np.ascontiguousarray(indexer, dtype=np.int64)
    valid_blocks = np.ascontiguousarray(valid_blocks, dtype=np.int64)
    mapping = np.ascontiguousarray(vec_batch_mapping, dtype=np.int32)
    cube_sdf = np.ascontiguousarray(cube_sdf, dtype=np.float32)
    cube_std = np.ascontiguousarray(cube_std, dtype=np.float32)
    assert max_n_triangles > 0
    r = cube_sdf.shape[1] // 2
    tri = np.empty((max_n_triangles, 3, 3), np.float32)
    fid = np.empty((max_n_triangles,), np.int64)
    std = np.empty((max_n_triangles, 3), np.float32)
    n = _lib().dif_oracle_marching_cubes(
        indexer.ctypes.data, valid_blocks.ctypes.data, valid_blocks.shape[0], mapping.ctypes.data, mapping.shape[0],
        cube_sdf.ctypes.data, cube_std.ctypes.data, r, int(n_xyz[0]), int(n_xyz[1]), int(n_xyz[2]),
        float(max_std), int(max_n_triangles), tri.ctypes.data, fid.ctypes.data, std.ctypes.data)
    n = min(int(n), max_n_triangles)
    return tri[:n], fid[:n], std[:n]
