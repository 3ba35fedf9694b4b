"""Golden vectors for the mesh-cache merge: the keep mask of the reference's OWN `_get_valid_idx`
(/root/reference/pytorch/system/map.py:20-26, numba-jitted, executed here through oracle/ref_shim.py) on the seeded id
streams that tests/test_gpu_parity.py::test_device_mesh_cache_merge_matches_reference_host_merge replays.  Build container only.

    python tests/golden/make_golden_merge.py      ->  tests/golden/ref_host_merge.npz

Every step's new ids contain the largest id (49 999), so `np.searchsorted` never returns len(query) - for a cached id above
max(query) the reference indexes one past the end of the array (undefined under numba); that case is outside the fixture and
our merge keeps such rows (the id is not among the new ones), which is what the reference intends.
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import dif_oracle as O               # noqa: E402
from oracle import ref_shim                       # noqa: E402

STEPS = [(5000, 400), (3000, 800), (0, 1), (7777, 50_000), (1, 3)]
SENTINEL = 49_999


def id_stream(seed=3):
    """The per-step new PLIVox ids (shared with the GPU test)."""
    rng = np.random.default_rng(seed)
    for n_new, id_hi in STEPS:
        fid = np.sort(rng.integers(0, id_hi, n_new)).astype(np.int64)
        rng.shuffle(fid)
        yield np.concatenate([fid, np.array([SENTINEL], np.int64)])


def main():
    ref = ref_shim.load_reference()
    out = {}
    cache_ids = None
    for step, fid in enumerate(id_stream()):
        if cache_ids is None:
            keep = np.zeros(0, bool)
            cache_ids = fid
        else:
            p = np.sort(np.unique(fid))                                  # map.py:708
            keep = np.asarray(ref.map._get_valid_idx(cache_ids, p))      # map.py:709, the reference's function, executed
            assert np.array_equal(keep, O.host_cache_keep_mask(cache_ids, fid)), "oracle restatement differs from the reference"
            cache_ids = np.concatenate([cache_ids[keep], fid])
        out[f"s{step}.new_ids"] = fid
        out[f"s{step}.keep"] = np.packbits(keep)
        out[f"s{step}.n_cache_after"] = np.int64(cache_ids.shape[0])
    np.savez_compressed(Path(__file__).resolve().parent / "ref_host_merge.npz", **out)
    print("ok", {k: v.shape for k, v in out.items() if k.endswith("new_ids")}, int(cache_ids.shape[0]))


if __name__ == "__main__":
    main()
