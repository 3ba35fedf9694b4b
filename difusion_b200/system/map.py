"""``DenseIndexedMap`` - host-side mirror of the reference's sparse PLIVox latent map (reference system/map.py:158-723)
for the fusion hot path: ``integrate_keyframe``, ``get_sdf``, ``extract_mesh``, ``save``/``load``, ``allocate_block`` and
the public state attributes.  Same constructor, method names, argument meaning and return values, so it drops into the
reference's SLAM loop (main.py:71-94); every computation is a call into libdifusion_b200.so (C ABI,
include/difusion_b200.h) on the current CUDA stream.  There is no torch/CPU fallback path.

Deliberate differences (documented in DESIGN.md):
  * latent buffers are physically pre-sized (power-of-two, grown by doubling) while the public ``latent_vecs`` /
    ``latent_vecs_pos`` / ``voxel_obs_count`` / ``voxel_optimized`` views keep the reference's capacity rule
    (smallest power of two >= n_occupied, map.py:263-281) so shapes match the reference exactly;
  * ``n_occupied`` lives on the device; the host value is fetched lazily (the reference syncs ~25 times per integrate);
  * points outside the grid are dropped, counted (``n_frames_with_dropped_points``, ``last_integrate_stats["flags"] & 1``) and
    logged instead of indexing out of range (map.py:313 "will not check index overflow");
  * the latent-optimisation branch (do_optimize, map.py:456-516, never enabled by main.py:85-86) is built for the synchronous mode only
    (``_optimize_latents`` / ``optimize_latent_rows`` over ``dif_latent_grad``); ``async_optimize`` (a forked process) raises.
"""
from __future__ import annotations

import argparse
import ctypes
import logging
import threading
from pathlib import Path

import numpy as np
import torch

from .. import _lib
from ..network import utility as net_util
from . import ext as _ext

LATENT_DIM = 29


def _as_f32(t: torch.Tensor) -> torch.Tensor:
    """Contiguous fp32 view of a caller tensor without the three no-op dispatches of .detach().contiguous().float()."""
    if t.dtype == torch.float32 and t.is_contiguous() and not t.requires_grad:
        return t
    return t.detach().contiguous().float()


def _next_pow2(v: int) -> int:
    p = 1
    while p < v:
        p *= 2
    return p


class MeshExtractCache:                                  # reference map.py:116-133
    """The reference keeps the cached mesh in host numpy arrays and merges on the host (map.py:698-714).  Here the cache lives
    on the device (``d_vertices`` (T,3,3) world coordinates, ``d_flatten_id`` (T,), ``d_std`` (T,3)) and is merged by
    ``dif_mesh_cache_merge``; the reference's attribute names ``vertices`` / ``vertices_flatten_id`` / ``vertices_std`` remain
    readable and download on demand."""

    def __init__(self, owner):
        self._owner = owner
        self.d_vertices = None
        self.d_flatten_id = None
        self.d_std = None
        self._host = None
        self.device = owner.device

    def _set(self, tri, fid, std):
        self.d_vertices, self.d_flatten_id, self.d_std, self._host = tri, fid, std, None

    def _download(self):
        if self.d_vertices is None:
            return None
        if self._host is None:
            self._host = (self.d_vertices.cpu().numpy(), self.d_flatten_id.cpu().numpy(), self.d_std.cpu().numpy())
        return self._host

    vertices = property(lambda self: None if self.d_vertices is None else self._download()[0])
    vertices_flatten_id = property(lambda self: None if self.d_vertices is None else self._download()[1])
    vertices_std = property(lambda self: None if self.d_vertices is None else self._download()[2])

    @property
    def updated_vec_id(self) -> torch.Tensor:
        """Sorted unique ids of PLIVoxes fused since the last extraction (map.py:303-308); kept as per-slot flags on the device."""
        o = self._owner
        return torch.nonzero(o._dirty[:o.n_occupied]).flatten()

    def clear_updated_vec(self):
        self._owner._dirty.zero_()

    def clear_all(self):
        self._set(None, None, None)
        self.clear_updated_vec()


class TriangleMesh:
    """What extract_mesh returns (the reference builds an open3d TriangleMesh, map.py:521-543; open3d is a GUI
    dependency outside the hot path).  Fields mirror what the reference fills in.  Built from device tensors, it downloads
    lazily: ``n_triangles`` / ``has_triangles()`` cost nothing, ``vertices`` / ``triangles`` / ``vertex_std`` copy on first use;
    ``d_vertices`` (T,3,3) / ``d_std`` (T,3) are the device views for consumers that stay on the GPU."""

    def __init__(self, vertices, vertex_std):
        self.d_vertices, self.d_std = (vertices, vertex_std) if torch.is_tensor(vertices) else (None, None)
        self._v, self._s = (None, None) if torch.is_tensor(vertices) else (vertices, vertex_std)
        self.n_triangles = int(vertices.shape[0])
        self._cooked = None

    def _host(self):
        if self._cooked is None:
            v = self.d_vertices.cpu().numpy() if self._v is None else self._v
            s = self.d_std.cpu().numpy() if self._s is None else self._s
            v = v.reshape(-1, 3).astype(float)
            self._cooked = (v, np.arange(v.shape[0], dtype=np.int32).reshape(-1, 3), s.reshape(-1).astype(float))
        return self._cooked

    vertices = property(lambda self: self._host()[0])
    triangles = property(lambda self: self._host()[1])
    vertex_std = property(lambda self: self._host()[2])

    def has_triangles(self):
        return self.n_triangles > 0


class DenseIndexedMap:
    STATUS_CONF_BIT = 1 << 0
    STATUS_SURF_BIT = 1 << 1

    def __init__(self, model, args: argparse.Namespace, latent_dim: int, device: torch.device, enable_async: bool = False,
                 optimization_device: torch.device = None, initial_capacity: int = 1 << 16, shard=None, shard_block_log2: int = 4,
                 initial_rows: int = None):
        device = torch.device(device)
        if device.type != "cuda":
            raise _lib.DifusionLibraryError("DenseIndexedMap needs a CUDA device: difusion_b200 has no CPU fallback")
        if device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        if latent_dim != LATENT_DIM:
            raise ValueError(f"the kernels are built for latent_dim={LATENT_DIM} (ckpt/default/hyper.json)")
        self._L = _lib.lib()
        self.model = model
        self.model.eval()
        self._prep = net_util.prepared_for(model, device)
        self.voxel_size = args.voxel_size
        self.n_xyz = np.ceil((np.asarray(args.bound_max) - np.asarray(args.bound_min)) / args.voxel_size).astype(int).tolist()
        logging.info(f"Map size Nx = {self.n_xyz[0]}, Ny = {self.n_xyz[1]}, Nz = {self.n_xyz[2]}")
        self.args = args
        self.bound_min = torch.tensor(args.bound_min, device=device).float()
        self.bound_max = self.bound_min + self.voxel_size * torch.tensor(self.n_xyz, device=device)
        self.latent_dim = latent_dim
        self.device = device
        self.extract_mesh_std_range = None
        # `tensor / python_float` (map.py:367,565): "true" = IEEE division (the reference on CPU tensors, the committed fixtures),
        # "reciprocal" = multiply by the fp32 reciprocal (what torch does on CUDA tensors: bit-identical cells to the reference run on a GPU)
        mode = getattr(args, "scalar_division", "true")
        if mode not in ("true", "reciprocal"):
            raise ValueError(f"scalar_division must be 'true' or 'reciprocal', got {mode!r}")
        self._division_mode = 1 if mode == "reciprocal" else 0
        self._n_cells = int(np.prod(self.n_xyz))
        if self._n_cells >= 2 ** 31:
            raise ValueError("grid too large for 32-bit linear ids")

        with torch.cuda.device(device):
            self._indexer = torch.full((self._n_cells,), -1, dtype=torch.long, device=device)
            self._n_occ_dev = torch.zeros(1, dtype=torch.int32, device=device)
            self._stats_dev = torch.zeros(_lib.DIF_STAT_COUNT, dtype=torch.int32, device=device)
        # integrate results come back through a small ring of pinned buffers; the host only waits when it needs a value
        self._stats_ring = [(torch.zeros(_lib.DIF_STAT_COUNT, dtype=torch.int32).pin_memory(), torch.cuda.Event()) for _ in range(8)]
        self._stats_inflight = []            # [(ring index, max new slots of that call)]
        self._stats_next = 0
        self._stats_last = [0] * _lib.DIF_STAT_COUNT
        self._n_occ_host = 0
        self._pending_error = None
        self._n_rows_host = 0
        self._reset_epoch = 0
        self.n_frames_with_dropped_points = 0
        # hash-sharded map (difusion_b200/shard.py): `shard` carries rank / world; the latent table then holds only the rows this
        # rank owns or keeps in its halo, addressed through _row_of (slot -> local row)
        self._shard_rank, self._shard_world, self._xchg = 0, 1, None
        self._shard_k, self._row_of, self._n_rows_dev, self._row_cap = 0, None, None, 0
        self._cap_phys = 0
        self._latent = self._pos = self._obs = self._optimized = self._dirty = None
        if shard is not None and shard.world > 1:
            self._shard_rank, self._shard_world, self._shard_k = shard.rank, shard.world, int(shard_block_log2)
            with torch.cuda.device(device):
                self._row_of = torch.full((1,), -1, dtype=torch.int32, device=device)
                self._n_rows_dev = torch.zeros(1, dtype=torch.int32, device=device)
                self._xchg = torch.empty(1 << 16, dtype=torch.int32, device=device)
            self._grow_rows(max(1, int(initial_rows or initial_capacity)))
        self._persist = None
        self._scratch = None
        self._scratch_points = 0
        self._mesh_persist = None
        self._mesh_persist_cap = 0
        self._cache_persist = None
        self._icp_scratch = None
        self._view_key, self._view_obj = None, None
        self._icp_out = None
        self._grow(max(1, _next_pow2(int(initial_capacity))))

        self.modifying_lock = threading.Lock()
        self.meshing_thread = None
        self.meshing_thread_id = -1
        self.meshing_stream = torch.cuda.Stream(device=device)
        self.mesh_cache = MeshExtractCache(self)
        self.last_integrate_stats = None

    # ------------------------------------------------------------------ state (reference cold_vars, map.py:199-211)
    def _grow(self, new_cap: int):
        """Double the per-slot arrays.  On a sharded map (``_row_of`` set, difusion_b200/shard.py) the latent table is indexed by
        LOCAL row, not by slot, and grows separately (``_grow_rows``)."""
        dev = self.device
        sharded_rows = self._row_of is not None
        with torch.cuda.device(dev):
            pos = torch.full((new_cap,), -1, dtype=torch.long, device=dev)
            obs = torch.zeros((new_cap,), dtype=torch.float32, device=dev)
            opt = torch.zeros((new_cap,), dtype=torch.bool, device=dev)
            dirty = torch.zeros((new_cap,), dtype=torch.uint8, device=dev)
            if self._cap_phys:
                c = self._cap_phys
                pos[:c], obs[:c], opt[:c], dirty[:c] = self._pos, self._obs, self._optimized, self._dirty
            self._pos, self._obs, self._optimized, self._dirty = pos, obs, opt, dirty
            if sharded_rows:
                row_of = torch.full((new_cap,), -1, dtype=torch.int32, device=dev)
                row_of[:self._cap_phys] = self._row_of[:self._cap_phys]
                self._row_of = row_of
            else:
                lat = torch.zeros((new_cap, _lib.LATENT_ROW_FLOATS), dtype=torch.float32, device=dev)   # 128-byte rows, columns 29..31 stay zero
                if self._cap_phys:
                    lat[:self._cap_phys] = self._latent
                self._latent = lat
                self._row_cap = new_cap
            self._cap_phys = new_cap
            self._persist = torch.zeros(self._L.dif_integrate_persist_bytes(self._n_cells, new_cap), dtype=torch.uint8, device=dev)

    def _grow_rows(self, new_rows: int):
        """Sharded map: grow the local latent table (rows = PLIVoxes this rank owns or keeps in its halo)."""
        lat = torch.zeros((new_rows, _lib.LATENT_ROW_FLOATS), dtype=torch.float32, device=self.device)
        if self._latent is not None:
            lat[:self._latent.size(0)] = self._latent
        self._latent, self._row_cap = lat, new_rows

    def _retire_stats(self, block: bool):
        """Consume finished integrate results (all of them when block=True).  Nothing is ever raised from the non-blocking path:
        which later call would see a finished event depends on timing, and a rank of a sharded map that raised alone would
        leave the others inside a collective.  Out-of-grid points are dropped, counted and logged (the reference indexes out of
        range there, map.py:313); an exhausted capacity (internal sizing error) is raised at the next BLOCKING point."""
        while self._stats_inflight:
            buf, ev = self._stats_ring[self._stats_inflight[0][0]]
            if block:
                ev.synchronize()
            elif not ev.query():
                break
            self._stats_inflight.pop(0)
            self._consume_stats(buf.tolist())
        if block and self._pending_error is not None:
            err, self._pending_error = self._pending_error, None
            raise err

    def _consume_stats(self, st):
        self._stats_last = st
        self._n_occ_host = st[_lib.STAT_N_OCCUPIED]
        self.last_integrate_stats = dict(n_kept=st[0], n_new=st[1], n_samples=st[2], n_updated=st[3], n_occupied=st[4],
                                         flags=st[5], n_focused=st[6])
        self._n_rows_host = st[_lib.STAT_N_ROWS]
        if st[_lib.STAT_FLAGS] & 6:
            self._pending_error = RuntimeError("PLIVox capacity exhausted inside dif_integrate (internal sizing error)")
        if st[_lib.STAT_FLAGS] & 1:
            self.n_frames_with_dropped_points += 1
            if self.n_frames_with_dropped_points in (1, 10, 100) or self.n_frames_with_dropped_points % 1000 == 0:
                logging.warning("integrate_keyframe: surface points outside the map bounds were dropped (%d frames so far; the "
                                "reference indexes out of range here, map.py:313)", self.n_frames_with_dropped_points)

    def reset(self):
        """Back to the empty map IN PLACE: every buffer keeps its address, so captured launch sequences (FramePipeline) stay
        valid.  The per-frame scratch is self-cleaning and therefore already zero."""
        torch.cuda.synchronize(self.device)
        self._stats_inflight.clear()
        self._indexer.fill_(-1)
        self._latent.zero_(); self._pos.fill_(-1); self._obs.zero_(); self._optimized.zero_(); self._dirty.zero_()
        self._n_occ_dev.zero_()
        if self._row_of is not None:
            self._row_of.fill_(-1); self._n_rows_dev.zero_()
        self._n_occ_host = 0
        self._n_rows_host = 0
        self._pending_error = None
        self._reset_epoch += 1
        self.last_integrate_stats = None
        self.mesh_cache.clear_all()
        torch.cuda.synchronize(self.device)

    def frame_pipeline(self, max_points: int, **kw):
        """One-launch-per-frame path (tracker linearisation + integrate as a replayable CUDA graph), see system/frame.py."""
        from .frame import FramePipeline
        return FramePipeline(self, max_points, **kw)

    def _sync_stats(self):
        self._retire_stats(block=True)

    @property
    def n_occupied(self) -> int:
        self._sync_stats()
        return self._n_occ_host

    @n_occupied.setter
    def n_occupied(self, v: int):
        self._sync_stats()
        self._n_occ_host = int(v)
        self._n_occ_dev.fill_(int(v))

    def _cap_ref(self) -> int:
        return max(1, _next_pow2(self.n_occupied))         # the reference's doubling rule, map.py:266-268

    indexer = property(lambda self: self._indexer)
    latent_vecs = property(lambda self: self._latent[:self._cap_ref(), :LATENT_DIM])        # the reference's (capacity, 29) tensor: a strided view
    latent_vecs_pos = property(lambda self: self._pos[:self._cap_ref()])
    voxel_obs_count = property(lambda self: self._obs[:self._cap_ref()])
    voxel_optimized = property(lambda self: self._optimized[:self._cap_ref()])

    @property
    def cold_vars(self) -> dict:
        return {"n_occupied": self.n_occupied, "indexer": self.indexer, "latent_vecs": self.latent_vecs,
                "latent_vecs_pos": self.latent_vecs_pos, "voxel_obs_count": self.voxel_obs_count,
                "voxel_optimized": self.voxel_optimized}

    def save(self, path):                                   # map.py:239-243, same on-disk dict
        with Path(path).open("wb") as f:
            torch.save({k: (v.clone() if torch.is_tensor(v) else v) for k, v in self.cold_vars.items()}, f)

    def load(self, path):                                   # map.py:245-249
        with Path(path).open("rb") as f:
            cv = torch.load(f, map_location=self.device)
        n = int(cv["n_occupied"])
        assert cv["indexer"].numel() == self._n_cells, "map bounds differ from the saved map"
        if n > self._cap_phys:
            self._grow(_next_pow2(n))
        c = cv["latent_vecs"].size(0)
        self._indexer.copy_(cv["indexer"])
        self._latent.zero_(); self._pos.fill_(-1); self._obs.zero_(); self._optimized.zero_(); self._dirty.zero_()
        self._latent[:c, :LATENT_DIM], self._pos[:c], self._obs[:c], self._optimized[:c] = cv["latent_vecs"], cv["latent_vecs_pos"], \
            cv["voxel_obs_count"], cv["voxel_optimized"]
        self.n_occupied = n

    def _view(self) -> _lib.MapView:
        key = (self._indexer.data_ptr(), self._latent.data_ptr(), self._cap_phys, self._shard_rank, self._shard_world, self._row_cap,
               self._xchg.data_ptr() if self._xchg is not None else 0)
        if self._view_key == key:
            return self._view_obj
        a = self.args
        v = _lib.MapView()
        v.indexer, v.latent_vecs, v.latent_vecs_pos = self._indexer.data_ptr(), self._latent.data_ptr(), self._pos.data_ptr()
        v.voxel_obs_count, v.slot_dirty, v.n_occupied = self._obs.data_ptr(), self._dirty.data_ptr(), self._n_occ_dev.data_ptr()
        v.capacity = self._cap_phys
        v.nx, v.ny, v.nz = self.n_xyz
        bm = np.asarray(a.bound_min, dtype=np.float32)
        v.bound_min[0], v.bound_min[1], v.bound_min[2] = float(bm[0]), float(bm[1]), float(bm[2])
        v.voxel_size = float(np.float32(a.voxel_size))
        v.prune_min_vox_obs = int(a.prune_min_vox_obs)
        v.ignore_count_th = float(a.ignore_count_th)
        v.encoder_count_th = float(a.encoder_count_th)
        v.shard_rank, v.shard_world = self._shard_rank, self._shard_world
        v.xchg_slots = self._xchg.data_ptr() if self._xchg is not None else None
        v.latent_stride = _lib.LATENT_ROW_FLOATS
        v.shard_block_log2 = self._shard_k
        v.row_of_slot = self._row_of.data_ptr() if self._row_of is not None else None
        v.n_rows = self._n_rows_dev.data_ptr() if self._n_rows_dev is not None else None
        v.row_capacity = self._row_cap
        v.scalar_division_mode = self._division_mode
        self._view_key, self._view_obj = key, v
        return v

    # ------------------------------------------------------------------ addressing helpers (map.py:287-319)
    def _linearize_id(self, xyz: torch.Tensor):
        return xyz[:, 2] + self.n_xyz[-1] * xyz[:, 1] + (self.n_xyz[-1] * self.n_xyz[-2]) * xyz[:, 0]

    def _unlinearize_id(self, idx: torch.Tensor):
        return torch.stack([idx // (self.n_xyz[1] * self.n_xyz[2]), (idx // self.n_xyz[2]) % self.n_xyz[1], idx % self.n_xyz[2]], dim=-1)

    def allocate_block(self, idx: torch.Tensor):
        """map.py:310-319: append slots for the given (N,3) or (N,) cell ids in the given order (host-driven, rarely used)."""
        if idx.ndimension() == 2 and idx.size(1) == 3:
            idx = self._linearize_id(idx)
        n0, k = self.n_occupied, idx.size(0)
        if n0 + k > self._cap_phys:
            self._grow(_next_pow2(n0 + k))
        new_id = torch.arange(n0, n0 + k, device=self.device, dtype=torch.long)
        self._pos[new_id] = idx
        self._indexer[idx] = new_id
        self.n_occupied = n0 + k

    # ------------------------------------------------------------------ integrate (map.py:340-519)
    def integrate_keyframe(self, surface_xyz: torch.Tensor, surface_normal: torch.Tensor, do_optimize: bool = False,
                           async_optimize: bool = False):
        assert surface_xyz.device == surface_normal.device == self.device, \
            f"Device of map {self.device} and input observation {surface_xyz.device, surface_normal.device} must be the same."
        if do_optimize and async_optimize:
            raise NotImplementedError("async_optimize forks a second process with its own decoder copy (map.py:28-78); only the "
                                      "synchronous latent optimisation is built")
        xyz, nrm = _as_f32(surface_xyz), _as_f32(surface_normal)
        n = xyz.size(0)
        with self.modifying_lock:
            # capacity: a call allocates at most 7 cells per point (own cell + 6 face neighbours); calls still in flight
            # are accounted with the same bound, so the host never has to wait for a count in the steady state
            self._retire_stats(block=False)
            worst = min(7 * n, self._n_cells)
            if len(self._stats_inflight) >= len(self._stats_ring) - 1 or \
                    self._n_occ_host + sum(w for _, w in self._stats_inflight) + worst > self._cap_phys:
                self._retire_stats(block=True)
                if self._n_occ_host + worst > self._cap_phys:
                    self._grow(_next_pow2(self._n_occ_host + worst))
            if self._row_of is not None:                     # sharded map: the local latent table follows the same worst-case rule
                pending = worst * (len(self._stats_inflight) + 1)
                if self._n_rows_host + pending > self._row_cap:
                    self._retire_stats(block=True)
                    if self._n_rows_host + worst > self._row_cap:
                        self._grow_rows(max(2 * self._row_cap, self._n_rows_host + 2 * worst))
            if n > self._scratch_points:
                self._scratch_points = max(n, 1 << 15)
                self._scratch = torch.empty(self._L.dif_integrate_scratch_bytes(self._scratch_points), dtype=torch.uint8, device=self.device)
            prune = self.args.prune_min_vox_obs > 0
            unq = torch.empty(n, dtype=torch.uint8, device=self.device) if prune else None
            view = self._view()
            st = _lib.stream_ptr(self.device)
            _lib.check(self._L.dif_integrate(ctypes.byref(view), self._prep.encoder.data_ptr(), xyz.data_ptr(), nrm.data_ptr(), n, None,
                                             _lib.ptr(unq), self._persist.data_ptr(), self._persist.numel(), self._scratch.data_ptr(),
                                             self._scratch.numel(), self._stats_dev.data_ptr(), st), "dif_integrate")
            buf, ev = self._stats_ring[self._stats_next]
            buf.copy_(self._stats_dev, non_blocking=True)
            ev.record(torch.cuda.current_stream(self.device))
            self._stats_inflight.append((self._stats_next, worst))
            self._stats_next = (self._stats_next + 1) % len(self._stats_ring)
            mask = unq.view(torch.bool) if prune else None
            if do_optimize and getattr(self.args, "optim_n_iters", 0) > 0:       # map.py:456 (the optimise process is never busy in sync mode)
                self._optimize_latents(xyz, nrm, mask)
        return mask

    # ------------------------------------------------------------------ latent optimisation (map.py:453-516, 80-117)
    optim_noise_fn = None       # callable(n, device) -> (n,) f32 standard-normal samples; None = torch.randn (map.py:486)

    def _normalize(self, xyz: torch.Tensor) -> torch.Tensor:
        """(xyz - bound_min) / voxel_size with the map's scalar-division rule (see scalar_division in __init__)."""
        z = xyz - self.bound_min.unsqueeze(0)
        vs = torch.tensor(self.voxel_size, dtype=torch.float32, device=self.device)
        return z * (1.0 / vs) if self._division_mode == 1 else z / vs          # tensor / tensor is a true division on CUDA too

    def _expand_flatten_id(self, base: torch.Tensor) -> torch.Tensor:
        """map.py:545-557 with ensure_valid=False: the cells and their (grid-clamped) 6 face neighbours, sorted unique."""
        pos = self._unlinearize_id(base)
        out = [base]
        for off in ((-1, 0, 0), (1, 0, 0), (0, -1, 0), (0, 1, 0), (0, 0, -1), (0, 0, 1)):
            q = pos + torch.tensor([off], device=self.device)
            for dim in range(3):
                q[:, dim].clamp_(0, self.n_xyz[dim] - 1)
            out.append(self._linearize_id(q))
        return torch.unique(torch.cat(out))

    def _optimize_latents(self, xyz: torch.Tensor, nrm: torch.Tensor, mask):
        """Step 3 of integrate_keyframe (map.py:453-516) in synchronous mode: PLIVoxes whose confidence reached encoder_count_th and
        that were not optimised yet get their latents refined by `optim_n_iters` Adam steps on the decoder's log-likelihood of
        noisy samples around the frame's surface points.  Index work as torch ops (this branch is disabled in the shipped loop,
        main.py:85-86); the decoder forward + backward of every iteration is dif_latent_grad."""
        a = self.args
        n_occ = self.n_occupied
        cap = self._cap_ref()
        obs, pos, opt = self._obs[:cap], self._pos[:cap], self._optimized[:cap]
        optim_pos = pos[torch.logical_and(obs >= a.encoder_count_th, ~opt)]
        optim_pos = optim_pos[optim_pos > 0]                                   # (sic, map.py:459: '> 0', cell 0 never qualifies)
        if optim_pos.size(0) == 0:
            return
        p_norm = self._normalize(xyz)
        if mask is not None:
            p_norm, nrm = p_norm[mask], nrm[mask]
        grid_id = self._linearize_id(torch.ceil(p_norm).long() - 1)
        status = torch.zeros(self._n_cells, dtype=torch.uint8, device=self.device)
        status[optim_pos] = 1
        focus = torch.zeros(self._n_cells, dtype=torch.bool, device=self.device)
        focus[self._expand_flatten_id(optim_pos)] = True
        fm = focus[grid_id]                                                     # get_pruned_surface (map.py:389-399)
        p_norm, nrm = p_norm[fm], nrm[fm]
        noise_fn = self.optim_noise_fn or (lambda n, dev: torch.randn(n, device=dev, dtype=torch.float32))
        inds, rels, sdfs = [], [], []
        for off in ((-0.5, -0.5, -0.5), (-0.5, -0.5, 0.5), (-0.5, 0.5, -0.5), (-0.5, 0.5, 0.5),
                    (0.5, -0.5, -0.5), (0.5, -0.5, 0.5), (0.5, 0.5, -0.5), (0.5, 0.5, 0.5)):
            gid = torch.ceil(p_norm + torch.tensor(off, device=self.device, dtype=torch.float32)) - 1
            for dim in range(3):
                gid[:, dim].clamp_(0, self.n_xyz[dim] - 1)
            rel = p_norm - gid - 0.5
            lin = self._linearize_id(gid.long())
            m = status[lin] >= self.STATUS_CONF_BIT
            cur_rel, cur_n = rel[m], nrm[m]
            cur_sdf = noise_fn(cur_rel.size(0), self.device) * 0.05
            inds.append(self._indexer[lin][m])
            rels.append(cur_rel + cur_sdf.unsqueeze(-1) * cur_n)
            sdfs.append(cur_sdf)
        inds, rels, sdfs = torch.cat(inds), torch.cat(rels).contiguous(), torch.cat(sdfs).contiguous()
        if inds.numel() == 0:
            return
        uniq, inv = torch.unique(inds, return_inverse=True)
        new_vecs = self.optimize_latent_rows(self._latent[uniq, :LATENT_DIM].contiguous(), inv.contiguous(), sdfs, rels)
        self._latent[uniq, :LATENT_DIM] = new_vecs                             # _update_optimize_result_set(deintegrate_old=False), map.py:320-336
        self._dirty[uniq] = 1
        self._optimized[uniq] = True
        del n_occ

    def optimize_latent_rows(self, latent_vecs_unique: torch.Tensor, latent_id_inv_mapping: torch.Tensor, gathered_sdf: torch.Tensor,
                             gathered_relative_xyz: torch.Tensor) -> torch.Tensor:
        """OptimizeProcess.do_optimize (map.py:80-117): Adam (lr 1e-2, torch defaults) on the unique latent rows; loss = decoder
        log-likelihood / n_samples (+ code_reg_lambda * sum ||row|| / n_samples per forward_model chunk if code_regularization)."""
        a = self.args
        lat = latent_vecs_unique.detach().clone().float().contiguous()
        inv = latent_id_inv_mapping.long().contiguous()
        sdf, rel = _as_f32(gathered_sdf), _as_f32(gathered_relative_xyz)
        n = inv.size(0)
        n_chunks = max(1, -(-n // int(1.5e6)))                                 # forward_model(max_sample=1.5e6): loss_func runs per chunk
        m1, m2 = torch.zeros_like(lat), torch.zeros_like(lat)
        grad = torch.empty_like(lat)
        lr, b1, b2, eps = 1.0e-2, 0.9, 0.999, 1e-8
        st = _lib.stream_ptr(self.device)
        for it in range(1, int(a.optim_n_iters) + 1):
            grad.zero_()
            _lib.check(self._L.dif_latent_grad(self._prep.decoder.data_ptr(), lat.data_ptr(), inv.data_ptr(), rel.data_ptr(), sdf.data_ptr(),
                                               n, n, grad.data_ptr(), None, st), "dif_latent_grad")
            if getattr(a, "code_regularization", False):                       # d/d row of lambda * ||row|| / n_samples, once per chunk (sic)
                nrm_r = torch.norm(lat, dim=1, keepdim=True)
                grad += (float(a.code_reg_lambda) * n_chunks / n) * torch.where(nrm_r > 0, lat / nrm_r, torch.zeros_like(lat))
            m1.mul_(b1).add_(grad, alpha=1 - b1)                               # torch.optim.Adam, default betas / eps, no weight decay
            m2.mul_(b2).addcmul_(grad, grad, value=1 - b2)
            denom = (m2.sqrt() / (1 - b2 ** it) ** 0.5).add_(eps)
            lat.addcdiv_(m1, denom, value=-lr / (1 - b1 ** it))
        return lat

    # ------------------------------------------------------------------ get_sdf (map.py:559-579)
    def get_sdf(self, xyz: torch.Tensor):
        """sdf (M,), std (M,), valid_mask (N,) - differentiable wrt ``xyz`` (the tracker calls autograd.grad on it)."""
        sdf, std, valid = _GetSdfFn.apply(xyz, self)
        return sdf, std, valid

    def _query(self, xyz: torch.Tensor):
        x = xyz.detach().contiguous().float()
        n = x.size(0)
        slot = torch.empty(n, dtype=torch.int32, device=self.device)
        rel = torch.empty((n, 3), dtype=torch.float32, device=self.device)
        nv = torch.empty(1, dtype=torch.int32, device=self.device)
        view = self._view()
        _lib.check(self._L.dif_map_query(ctypes.byref(view), x.data_ptr(), n, slot.data_ptr(), rel.data_ptr(), nv.data_ptr(),
                                         _lib.stream_ptr(self.device)), "dif_map_query")
        return slot, rel

    # ------------------------------------------------------------------ fused ICP linearisation (tracker.py:174-218)
    def icp_linearize(self, obs_xyz: torch.Tensor, R_last, t_last, R_delta, t_delta, huber_k: float = 5.0, want_grad: bool = True):
        """One launch: returns a pinned-host-bound device tensor out[44] (fp64): H[36], g[6], energy, M."""
        x = _as_f32(obs_xyz)
        n = x.size(0)
        if self._icp_scratch is None:
            self._icp_scratch = torch.zeros(self._L.dif_icp_scratch_bytes(n), dtype=torch.uint8, device=self.device)   # zero-filled once (ABI)
            self._icp_ring = torch.empty((16, 44), dtype=torch.float64, device=self.device)    # results of the last 16 calls stay valid
            self._icp_next = 0
        out = self._icp_ring[self._icp_next]
        self._icp_next = (self._icp_next + 1) % 16
        pose = np.empty(24, np.float32)
        pose[0:9], pose[9:12], pose[12:21], pose[21:24] = np.ravel(R_last), np.ravel(t_last), np.ravel(R_delta), np.ravel(t_delta)
        view = self._view()
        _lib.check(self._L.dif_icp_linearize(ctypes.byref(view), self._prep.decoder.data_ptr(), x.data_ptr(), n, pose.ctypes.data, None,
                                             float(huber_k) if huber_k else 0.0, int(want_grad), self._icp_scratch.data_ptr(),
                                             self._icp_scratch.numel(), out.data_ptr(), _lib.stream_ptr(self.device)), "dif_icp_linearize")
        return out

    # ------------------------------------------------------------------ mesh extraction (map.py:581-723)
    def _make_mesh_from_cache(self):
        if self.mesh_cache.d_vertices is None:
            return TriangleMesh(np.zeros((0, 3, 3), np.float32), np.zeros((0, 3), np.float32))
        return TriangleMesh(self.mesh_cache.d_vertices, self.mesh_cache.d_std)

    def _merge_into_cache(self, tri, fid, std):
        """map.py:698-714 on the device: voxel units -> world, drop cached triangles of PLIVoxes that produced new ones, append."""
        c, dev = self.mesh_cache, self.device
        n_cache = 0 if c.d_vertices is None else int(c.d_vertices.size(0))
        n_new = int(tri.size(0))
        need = self._L.dif_mesh_cache_scratch_bytes(self._n_cells, n_cache)
        if self._cache_persist is None or self._cache_persist.numel() < need:
            self._cache_persist = torch.zeros(self._L.dif_mesh_cache_scratch_bytes(self._n_cells, max(2 * n_cache, 1 << 20)),
                                              dtype=torch.uint8, device=dev)
        # fresh tensors (meshes handed out earlier keep their storage), sizes rounded so that the caching allocator finds a block
        n_out = -(-(n_cache + n_new) // (1 << 19)) * (1 << 19)
        o_tri = torch.empty((n_out, 3, 3), dtype=torch.float32, device=dev)
        o_id = torch.empty((n_out,), dtype=torch.long, device=dev)
        o_std = torch.empty((n_out, 3), dtype=torch.float32, device=dev)
        totals = torch.zeros(2, dtype=torch.long, device=dev)
        bm = (ctypes.c_float * 3)(*[float(np.float32(v)) for v in self.args.bound_min])
        _lib.check(self._L.dif_mesh_cache_merge(
            _lib.ptr(c.d_vertices), _lib.ptr(c.d_flatten_id), _lib.ptr(c.d_std), n_cache, _lib.ptr(tri), _lib.ptr(fid), _lib.ptr(std), n_new,
            float(np.float32(self.voxel_size)), bm, self._n_cells, o_tri.data_ptr(), o_id.data_ptr(), o_std.data_ptr(), totals.data_ptr(),
            self._cache_persist.data_ptr(), self._cache_persist.numel(), _lib.stream_ptr(dev)), "dif_mesh_cache_merge")
        total = int(totals[1].item())                        # host sync: the merged cache is sliced to its size
        c._set(o_tri[:total], o_id[:total], o_std[:total])

    def _snapshot(self):
        """What the reference's backup_vars are for (map.py:620-622): the tensors a mesh extraction reads, pinned by reference so
        that a concurrent capacity growth (which rebinds the attributes) cannot hand their memory back to the allocator, plus a
        COPY of the device view built from exactly these tensors."""
        v = _lib.MapView()
        ctypes.memmove(ctypes.byref(v), ctypes.byref(self._view()), ctypes.sizeof(v))
        return dict(view=v, cap=self._cap_phys, tensors=(self._indexer, self._latent, self._pos, self._obs, self._dirty, self._n_occ_dev),
                    indexer=self._indexer, pos=self._pos)

    def _ws(self, name: str, numel: int, dtype) -> torch.Tensor:
        """Grow-only workspace of the mesh path (cubes, decode scratch, sample lists, marching-cubes output): a full extraction at
        BASELINE config 4 needs ~2.5 GB of intermediates whose sizes change from call to call; asking the caching allocator for
        them every time made it split, miss and cudaMalloc/cudaFree (device-synchronising) - 8 to 50 ms of a 10 ms extraction.
        The returned view is valid until the next request for the same name."""
        ws = self.__dict__.setdefault("_mesh_ws", {})
        buf = ws.get(name)
        if buf is None or buf.numel() < numel or buf.dtype != dtype:
            buf = torch.empty(max(int(numel * 1.25), 1), dtype=dtype, device=self.device)
            ws[name] = buf
        return buf[:numel]

    def mesh_cubes(self, voxel_resolution: int, fast: bool = True, updated_vec_id: torch.Tensor = None, snap: dict = None):
        """Stages map.py:627-687 on the device: returns (focused_flatten_id (K,), vec_id_batch_mapping (cap,), high_sdf, high_std
        (B,2r,2r,2r) [sdf already negated], block_slots (B,), counts dict).  updated_vec_id None == all occupied PLIVoxes.
        snap: a _snapshot() taken under modifying_lock (asynchronous extraction); default = the live map.
        The returned tensors are views of the map's mesh workspace (`_ws`): valid until the next mesh_cubes / extract_mesh call."""
        dev, L = self.device, self._L
        st = _lib.stream_ptr(dev)
        snap = snap or self._snapshot()
        view, cap = snap["view"], snap["cap"]
        if self._mesh_persist is None or self._mesh_persist_cap != cap:
            self._mesh_persist = torch.zeros(L.dif_mesh_select_scratch_bytes(self._n_cells, cap), dtype=torch.uint8, device=dev)
            self._mesh_persist_cap = cap
        k_max = cap if updated_vec_id is None else int(updated_vec_id.numel())
        upd = None if updated_vec_id is None else updated_vec_id.to(torch.int32).contiguous()
        focused = self._ws("focused", max(k_max, 1), torch.long)
        block_slots = self._ws("block_slots", min(cap, 7 * max(k_max, 1)), torch.int32)
        mapping = self._ws("mapping", cap, torch.int32)
        counts = torch.zeros(2, dtype=torch.int32, device=dev)
        _lib.check(L.dif_mesh_select(ctypes.byref(view), _lib.ptr(upd), 0 if upd is None else upd.numel(), focused.data_ptr(),
                                     block_slots.data_ptr(), mapping.data_ptr(), counts.data_ptr(), self._mesh_persist.data_ptr(),
                                     self._mesh_persist.numel(), st), "dif_mesh_select")
        K, B = counts.tolist()                               # host sync: output shapes depend on it
        r = int(voxel_resolution)
        hr = 2 * r
        cube_sdf = self._ws("cube_sdf", B * hr ** 3, torch.float32).view(B, hr, hr, hr)
        cube_std = self._ws("cube_std", B * hr ** 3, torch.float32).view(B, hr, hr, hr)
        scratch = self._ws("decode_scratch", max(L.dif_mesh_decode_scratch_bytes(B, r), 256), torch.uint8)
        dcounts = torch.zeros(2, dtype=torch.int32, device=dev)
        _lib.check(L.dif_mesh_decode(ctypes.byref(view), self._prep.decoder.data_ptr(), block_slots.data_ptr(), B, r, int(bool(fast)),
                                     cube_sdf.data_ptr(), cube_std.data_ptr(), scratch.data_ptr(), scratch.numel(), dcounts.data_ptr(), st),
                   "dif_mesh_decode")
        return focused[:K], mapping, cube_sdf, cube_std, block_slots[:B], dcounts

    def _marching_cubes_ws(self, indexer, focused, mapping, cube_sdf, cube_std, max_n_triangles: int, max_std: float):
        """system.ext.marching_cubes_interp (mc.cpp:3-16) with its outputs in the mesh workspace instead of three fresh
        max_n_triangles-sized tensors per call; an upper bound of 5 triangles per sub-cube caps the workspace."""
        r = cube_sdf.size(1) // 2
        cap = int(min(int(max_n_triangles), 5 * r ** 3 * max(int(focused.numel()), 1)))
        tri = self._ws("mc_tri", cap * 9, torch.float32).view(cap, 3, 3)
        fid = self._ws("mc_fid", cap, torch.int64)
        std = self._ws("mc_std", cap * 3, torch.float32).view(cap, 3)
        count = torch.zeros(1, dtype=torch.int32, device=self.device)
        _lib.check(self._L.dif_marching_cubes(
            indexer.data_ptr(), int(self.n_xyz[0]), int(self.n_xyz[1]), int(self.n_xyz[2]), focused.data_ptr(), focused.size(0),
            mapping.data_ptr(), mapping.size(0), cube_sdf.data_ptr(), cube_std.data_ptr(), r, float(max_std),
            tri.data_ptr(), fid.data_ptr(), std.data_ptr(), cap, count.data_ptr(), _lib.stream_ptr(self.device)), "dif_marching_cubes")
        n = int(count.item())                      # the reference syncs here too (mc_interp_kernel.cu:367-369)
        if n > cap:
            import sys
            sys.stderr.write(f"Warning from marching cube: the max triangle number is too small {n} vs {max_n_triangles}\n")
            n = cap
        return tri[:n], fid[:n], std[:n]

    def extract_mesh(self, voxel_resolution: int, max_n_triangles: int, fast: bool = True, max_std: float = 2000.0,
                     extract_async: bool = False, no_cache: bool = False, interpolate: bool = True):
        if not interpolate:
            raise NotImplementedError("interpolate=False calls system.ext.marching_cubes, which the reference does not export "
                                      "(map.py:693 vs ext/__init__.py:19)")
        if self.meshing_thread is not None:                 # map.py:597-607
            if not self.meshing_thread.is_alive():
                self.meshing_thread = None
                self.meshing_thread_id = -1
                return self._make_mesh_from_cache()
            elif not extract_async:
                self.meshing_thread.join()
                return self._make_mesh_from_cache()
            else:
                return None

        with self.modifying_lock:                           # map.py:609-622
            if no_cache:
                updated = None
                self.mesh_cache.clear_all()
            else:
                updated = self.mesh_cache.updated_vec_id
                if updated.size(0) == 0:
                    return self._make_mesh_from_cache() if not extract_async else None
                self.mesh_cache.clear_updated_vec()
            self._sync_stats()
            if self._shard_world > 1:                       # hash-sharded map: every rank meshes the PLIVoxes it owns
                if updated is None:
                    updated = torch.arange(self.n_occupied, device=self.device)
                owned = self.owned_slots(updated)
                updated = self.decode_set(owned)            # owned + their 26 neighbours: same decode batch as a full extraction sees
            else:
                owned = None
            # the mesher reads THIS state even if integrate_keyframe grows (rebinds) the map meanwhile (map.py:620-622 backup_vars)
            snap = self._snapshot()
            if extract_async:
                for t in snap["tensors"]:
                    t.record_stream(self.meshing_stream)

        def do_meshing(res):
            torch.cuda.synchronize(self.device)
            with torch.cuda.stream(self.meshing_stream):
                focused, mapping, cube_sdf, cube_std, _, _ = self.mesh_cubes(res, fast, updated, snap)
                if owned is not None:                       # sharded map: triangles only for the PLIVoxes this rank owns
                    focused = snap["pos"][owned].contiguous()
                if cube_sdf.size(0) == 0 or focused.numel() == 0:
                    return
                vertices, vertices_flatten_id, vertices_std = self._marching_cubes_ws(
                    snap["indexer"], focused, mapping, cube_sdf, cube_std, max_n_triangles, max_std)
                self._merge_into_cache(vertices, vertices_flatten_id, vertices_std)

        if extract_async:
            self.meshing_thread = threading.Thread(target=do_meshing, args=(voxel_resolution,), daemon=True)
            self.meshing_thread.start()
            self.meshing_thread_id = self.meshing_thread.ident
            return None
        do_meshing(voxel_resolution)
        return self._make_mesh_from_cache()


class _GetSdfFn(torch.autograd.Function):
    """map.py:559-579 as one differentiable op: lookup kernel -> compaction -> decoder forward(+backward) kernel.
    backward: d/dxyz_world = (g_sdf * dsdf/drel + g_std * dstd/drel) / voxel_size, zero rows for invalid points."""

    @staticmethod
    def forward(ctx, xyz, m: DenseIndexedMap):
        slot, rel = m._query(xyz)
        valid = slot >= 0
        idx = torch.nonzero(valid).flatten()                # host sync, as in the reference's boolean-mask indexing
        M = idx.numel()
        assert M > 0                                        # the reference asserts on an empty batch (utility.py:84-85)
        rows = slot[idx].contiguous()
        x = rel[idx].contiguous()
        dev = xyz.device
        sdf = torch.empty(M, dtype=torch.float32, device=dev)
        std = torch.empty(M, dtype=torch.float32, device=dev)
        need = xyz.requires_grad
        g = torch.empty((M, 3), dtype=torch.float32, device=dev) if need else None
        _lib.check(m._L.dif_decode(m._prep.decoder.data_ptr(), m._latent.data_ptr(), _lib.LATENT_ROW_FLOATS, rows.data_ptr(), x.data_ptr(), M, None, 1.0,
                                   sdf.data_ptr(), std.data_ptr(), _lib.ptr(g), None, _lib.stream_ptr(dev)), "dif_decode")
        ctx.m, ctx.n = m, xyz.size(0)
        ctx.save_for_backward(idx, rows, x, g if need else torch.empty(0, device=dev))
        ctx.mark_non_differentiable(valid)
        return sdf, std, valid

    @staticmethod
    def backward(ctx, g_sdf, g_std, _g_valid):
        idx, rows, x, dsdf = ctx.saved_tensors
        if dsdf.numel() == 0:
            return None, None
        m = ctx.m
        grad_rel = g_sdf.unsqueeze(-1) * dsdf
        if g_std is not None and bool((g_std != 0).any()):
            M = x.size(0)
            s0 = torch.empty(M, dtype=torch.float32, device=x.device); s1 = torch.empty_like(s0)
            g0 = torch.empty((M, 3), dtype=torch.float32, device=x.device); g1 = torch.empty_like(g0)
            _lib.check(m._L.dif_decode(m._prep.decoder.data_ptr(), m._latent.data_ptr(), _lib.LATENT_ROW_FLOATS, rows.data_ptr(), x.data_ptr(), M, None, 1.0,
                                       s0.data_ptr(), s1.data_ptr(), g0.data_ptr(), g1.data_ptr(), _lib.stream_ptr(x.device)), "dif_decode")
            grad_rel = grad_rel + g_std.unsqueeze(-1) * g1
        grad = torch.zeros((ctx.n, 3), dtype=torch.float32, device=x.device)
        grad[idx] = grad_rel / m.voxel_size
        return grad, None
