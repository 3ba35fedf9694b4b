"""ORACLE (test infrastructure only) - builds the UNMODIFIED reference CUDA extensions into oracle/_ref/.

The reference's native ops (pytorch/system/ext/__init__.py:15-44) are torch C++/CUDA extensions that it JIT-compiles with
torch.utils.cpp_extension.load.  This recipe compiles the very same source files, where they lie under /root/reference
(nothing is copied into the repository), for sm_100 with the container's nvcc, and leaves one pybind module per extension in
oracle/_ref/<name>/<name>.so.  oracle/_ref/ is git-ignored but travels to the GPU box with the snapshot, where
tests/test_ref_ext_gpu.py loads the modules and compares the product kernels against the reference's own kernels on the same
device inputs - that is what pins the marching-cubes / groupby_sum stages (the reference ships no golden vectors).

Only tests/, __graft_entry__.build() and bench.py's reference legs may touch this file or its outputs.

    python oracle/build_ref.py [marching_cubes indexing imgproc pcproc]

stage_python() additionally stages the reference's own Python for the integrate / track path (read by oracle/ref_gpu.py only).
"""
from __future__ import annotations

import importlib.util
import os
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
OUT = HERE / "_ref"
REF_EXT = Path("/root/reference/pytorch/system/ext")

# name -> sources, exactly the lists of the reference's ext/__init__.py:15-44
MODULES = {
    "marching_cubes": ["marching_cubes/mc.cpp", "marching_cubes/mc_interp_kernel.cu"],
    "indexing": ["indexing/indexing.cpp", "indexing/indexing.cu"],
    "imgproc": ["imgproc/imgproc.cu", "imgproc/imgproc.cpp", "imgproc/photometric.cu"],
    "pcproc": ["pcproc/pcproc.cpp", "pcproc/pcproc.cu", "pcproc/cuda_kdtree.cu"],
}


def so_path(name: str) -> Path:
    return OUT / name / f"{name}.so"


def available(name: str) -> bool:
    return so_path(name).exists()


def build(names=None, verbose: bool = False) -> list[Path]:
    """Compile the named reference extensions (default: all that are missing).  Needs /root/reference; a no-op elsewhere.
    Several missing modules are compiled side by side in child processes (each takes minutes: torch headers)."""
    names = list(names or MODULES)
    todo = [n for n in names if not available(n)]
    if not todo:
        return [so_path(n) for n in names]
    if not REF_EXT.exists():
        return [so_path(n) for n in names if available(n)]
    if len(todo) > 1:
        import subprocess
        procs = [subprocess.Popen([sys.executable, str(Path(__file__).resolve()), n], stdout=None if verbose else subprocess.DEVNULL,
                                  stderr=None if verbose else subprocess.DEVNULL) for n in todo]
        for p in procs:
            p.wait()
        return [so_path(n) for n in names if available(n)]
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0")
    os.environ.setdefault("MAX_JOBS", str(min(8, os.cpu_count() or 1)))
    from torch.utils.cpp_extension import load
    for n in todo:
        bdir = OUT / n
        bdir.mkdir(parents=True, exist_ok=True)
        load(name=n, sources=[str(REF_EXT / s) for s in MODULES[n]], build_directory=str(bdir), verbose=verbose,
             is_python_module=False)       # compile + link only; importing it needs libcuda
        for junk in list(bdir.glob("*.o")) + list(bdir.glob("*.d")):      # only the .so has to travel to the GPU box
            junk.unlink()
    return [so_path(n) for n in names if available(n)]


REF_PY = Path("/root/reference/pytorch")
PY_OUT = OUT / "pytorch"
# the reference's own Python on the integrate / track path and the shipped checkpoint (map.py, tracker.py and what they import)
PY_FILES = ["system/map.py", "system/tracker.py", "network/di_decoder.py", "network/di_encoder.py", "network/utility.py",
            "utils/exp_util.py", "utils/motion_util.py", "utils/pt_util.py", "dataset/production/__init__.py",
            "ckpt/default/hyper.json", "ckpt/default/model_300.pth.tar", "ckpt/default/encoder_300.pth.tar"]


def python_available() -> bool:
    return all((PY_OUT / f).exists() for f in PY_FILES)


def stage_python() -> bool:
    """Stage the UNMODIFIED reference Python of the hot path + checkpoint under oracle/_ref/pytorch/ (git-ignored, travels to the GPU
    box like the .so files) so that oracle/ref_gpu.py can run the reference's own torch-CUDA path there.  No-op without /root/reference."""
    if python_available() or not REF_PY.exists():
        return python_available()
    import shutil
    for f in PY_FILES:
        dst = PY_OUT / f
        dst.parent.mkdir(parents=True, exist_ok=True)
        shutil.copyfile(REF_PY / f, dst)
        os.chmod(dst, 0o644)
    return python_available()


def load_module(name: str):
    """Import oracle/_ref/<name>/<name>.so (a pybind module built by build()).  Needs a CUDA runtime; raises if not built."""
    import torch  # noqa: F401  (the module links against libtorch)
    p = so_path(name)
    if not p.exists():
        raise FileNotFoundError(f"{p} is missing: run `python oracle/build_ref.py {name}` in the build container")
    spec = importlib.util.spec_from_file_location(name, str(p))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    print("reference python staged:", stage_python())
    for p in build(sys.argv[1:] or None, verbose=True):
        print(p, p.stat().st_size)
