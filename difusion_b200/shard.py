"""Hash-sharded PLIVox map over several GPUs (new design: the reference is single-GPU; SURVEY 8e, BASELINE north_star).

* Every rank receives the same frame and runs the same integer index kernels, so ``indexer``, slot numbering,
  ``latent_vecs_pos`` and ``voxel_obs_count`` are replicated and stay bit-identical to the single-GPU map.
* Floating-point work is sharded by ``owner(cell) = splitmix64(linear id) % world``: the encoder MLP and the latent
  fusion of a PLIVox run only on its owner (``dif_map_view.shard_rank/shard_world``).
* One exchange per frame, ONE collective: each rank packs the (slot, latent row) pairs it owns and changed into a fixed-size
  buffer with a count header (dif_shard_pack), a single NCCL all-gather moves the buffers, dif_shard_unpack scatters the other
  ranks' rows into the local table - no host synchronisation, sizes never leave the device.  After it every rank holds every
  latent, so decoding / meshing need no further communication.  A rank that ever publishes more rows than the buffer holds
  raises a device flag; the host notices it at the next frame, doubles the buffer and re-synchronises all owned rows
  (variable-length path, also used by the CPU/gloo test).
* ICP linearisation: each rank processes a contiguous slice of the frame's points, one all-reduce of 44 doubles.
* Mesh extraction: each rank meshes the PLIVoxes it owns (neighbour cubes are decoded locally).

``ShardGroup`` is pure ``torch.distributed`` plumbing (NCCL on GPUs; the same code runs on CPU tensors with gloo, which is how
tests/test_shard_gloo.py exercises it).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

_M64 = (1 << 64) - 1


def owner_of_np(lin_ids, world: int) -> np.ndarray:
    """splitmix64(linear id) % world on the host (numpy uint64) - mirror of csrc/common.cuh shard_owner()."""
    with np.errstate(over="ignore"):
        z = np.asarray(lin_ids).astype(np.uint64) + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return (z % np.uint64(world)).astype(np.int64)


def owner_of(lin_ids: torch.Tensor, world: int) -> torch.Tensor:
    """The same function on a torch int64 tensor (any device): int64 arithmetic wraps, shifts are made logical by masking."""
    def lsr(v, k):
        return (v >> k) & ((1 << (64 - k)) - 1)

    def c(v):                     # python int -> wrapped int64 constant
        return v - (1 << 64) if v >= (1 << 63) else v
    z = lin_ids.to(torch.int64) + c(0x9E3779B97F4A7C15)
    z = (z ^ lsr(z, 30)) * c(0xBF58476D1CE4E5B9)
    z = (z ^ lsr(z, 27)) * c(0x94D049BB133111EB)
    z = z ^ lsr(z, 31)
    # unsigned modulo of a value stored in a signed int64
    r = torch.remainder(z, world)
    r = torch.where(z < 0, torch.remainder(r + ((1 << 64) % world), world), r)
    return r


class ShardGroup:
    def __init__(self, group=None):
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)

    def all_gather_rows(self, slots: torch.Tensor, rows: torch.Tensor):
        """Variable-length all-gather of (slot id, row) pairs.  Returns the concatenation over ranks (rank order)."""
        n = torch.tensor([slots.numel()], dtype=torch.int64, device=slots.device)
        counts = [torch.zeros_like(n) for _ in range(self.world)]
        dist.all_gather(counts, n, group=self.group)
        counts = [int(c.item()) for c in counts]
        m = max(counts)
        if m == 0:
            return slots[:0], rows[:0]
        pad_s = torch.zeros(m, dtype=slots.dtype, device=slots.device); pad_s[:slots.numel()] = slots
        pad_r = torch.zeros((m, rows.size(1)), dtype=rows.dtype, device=rows.device); pad_r[:rows.size(0)] = rows
        gs = [torch.empty_like(pad_s) for _ in range(self.world)]
        gr = [torch.empty_like(pad_r) for _ in range(self.world)]
        dist.all_gather(gs, pad_s, group=self.group)
        dist.all_gather(gr, pad_r, group=self.group)
        return torch.cat([g[:c] for g, c in zip(gs, counts)]), torch.cat([g[:c] for g, c in zip(gr, counts)])

    def all_gather_fixed(self, send: torch.Tensor, recv: torch.Tensor):
        """recv[rank] = send of that rank; one collective, sizes known statically."""
        dist.all_gather_into_tensor(recv, send, group=self.group)

    def all_reduce_sum(self, t: torch.Tensor):
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t


def combine_icp(out: torch.Tensor, group: ShardGroup) -> torch.Tensor:
    """Per-rank dif_icp_linearize outputs (already divided by the rank's valid count M) -> global H, g, energy, M."""
    raw = out.clone()
    raw[:43] *= out[43]
    group.all_reduce_sum(raw)
    M = raw[43]
    raw[:43] = torch.where(M > 0, raw[:43] / torch.clamp(M, min=1.0), torch.zeros_like(raw[:43]))
    return raw


def expand_26(indexer: torch.Tensor, pos: torch.Tensor, n_xyz, slots: torch.Tensor) -> torch.Tensor:
    """Slots of the occupied cells in the 3x3x3 neighbourhood (own cell included) of the given slots, sorted unique.
    Owner-wise meshing decodes this set so that every cube the marching-cubes blend of an OWNED PLIVox can touch
    (mc_interp_kernel.cu:103-181: +-1 per axis, diagonals included) is in the decode batch, exactly as in a full single-GPU extraction."""
    if slots.numel() == 0:
        return slots.to(torch.int64)
    nx, ny, nz = [int(v) for v in n_xyz]
    lin = pos[slots.long()]
    x, y, z = lin // (ny * nz), (lin // nz) % ny, lin % nz
    r = torch.arange(-1, 2, device=lin.device)
    ox, oy, oz = [t.reshape(1, -1) for t in torch.meshgrid(r, r, r, indexing="ij")]
    X, Y, Z = x[:, None] + ox, y[:, None] + oy, z[:, None] + oz
    ok = (X >= 0) & (X < nx) & (Y >= 0) & (Y < ny) & (Z >= 0) & (Z < nz)
    s = indexer[((X * ny + Y) * nz + Z)[ok]]
    return torch.unique(s[s >= 0])


def make_sharded_map(model, args, latent_dim, device, group: ShardGroup, **kw):
    from .system.map import DenseIndexedMap

    class ShardedMap(DenseIndexedMap):
        def __init__(self):
            super().__init__(model, args, latent_dim, device, **kw)
            self.shard = group
            self._shard_rank, self._shard_world = group.rank, group.world
            self._xchg = torch.empty(1 << 16, dtype=torch.int32, device=self.device)

        def _alloc_xchg(self, cap_rows: int):
            from . import _lib
            L = _lib.lib()
            self._xcap = cap_rows
            nfl = L.dif_shard_xchg_bytes(cap_rows) // 4
            self._xsend = torch.zeros(nfl, dtype=torch.float32, device=self.device)
            self._xrecv = torch.zeros(nfl * self.shard.world, dtype=torch.float32, device=self.device)
            self._xflag = torch.zeros(1, dtype=torch.int32, device=self.device)
            self._xflag_host = torch.zeros(1, dtype=torch.int32).pin_memory()
            self._xflag_ev = None

        def resync(self):
            """Publish ALL owned latent rows (variable-length path): recovery after an exchange-buffer overflow."""
            slots = self.owned_slots(torch.arange(self.n_occupied, device=self.device))
            slots_all, rows_all = self.shard.all_gather_rows(slots.int(), self._latent[slots])
            if slots_all.numel():
                self._latent.index_copy_(0, slots_all.long(), rows_all)

        def integrate_keyframe(self, surface_xyz, surface_normal, do_optimize=False, async_optimize=False):
            import ctypes
            from . import _lib
            L = _lib.lib()
            need = min(8 * surface_xyz.size(0), self._cap_phys) + 1
            if self._xchg.numel() < need:
                self._xchg = torch.empty(need, dtype=torch.int32, device=self.device)
            if getattr(self, "_xcap", 0) == 0:
                self._alloc_xchg(1 << 14)
            self._xframe = getattr(self, "_xframe", 0) + 1
            if self._xflag_ev is not None and self._xframe >= self._xflag_frame + 4:
                # did an exchange overflow?  Checked a fixed number of frames after the flag was copied (long complete, so the
                # wait is free) so that every rank takes the collective recovery path in the same frame; the flag itself is
                # identical on every rank because every rank sees every header.
                self._xflag_ev.synchronize()
                self._xflag_ev = None
                need = int(self._xflag_host[0])                                 # largest row count any rank tried to publish
                if need:
                    cap = self._xcap * 2
                    while cap < 2 * need:
                        cap *= 2
                    self._alloc_xchg(cap)
                    self.resync()
            mask = super().integrate_keyframe(surface_xyz, surface_normal, do_optimize, async_optimize)
            view, st = self._view(), _lib.stream_ptr(self.device)
            n_x = self._stats_dev[_lib.STAT_N_XCHG:]
            _lib.check(L.dif_shard_pack(ctypes.byref(view), n_x.data_ptr(), self._xcap, self._xsend.data_ptr(), st), "dif_shard_pack")
            self.shard.all_gather_fixed(self._xsend, self._xrecv)
            _lib.check(L.dif_shard_unpack(ctypes.byref(view), self._xrecv.data_ptr(), self.shard.world, self._xcap, self._xflag.data_ptr(), st),
                       "dif_shard_unpack")
            if self._xflag_ev is None:
                self._xflag_host.copy_(self._xflag, non_blocking=True)
                self._xflag_ev = torch.cuda.Event()
                self._xflag_ev.record(torch.cuda.current_stream(self.device))
                self._xflag_frame = self._xframe
            return mask

        @property
        def last_exchange(self):
            """Rows this rank published / all ranks published in the last frame (reads the headers: host sync; diagnostics only)."""
            per = self._xrecv.view(self.shard.world, -1)[:, 0].contiguous().view(torch.int32).tolist()
            return dict(rows_sent=per[self.shard.rank], rows_total=sum(per), capacity=self._xcap)

        def icp_linearize(self, obs_xyz, R_last, t_last, R_delta, t_delta, huber_k=5.0, want_grad=True):
            n = obs_xyz.size(0)
            lo, hi = n * self.shard.rank // self.shard.world, n * (self.shard.rank + 1) // self.shard.world
            out = super().icp_linearize(obs_xyz[lo:hi].contiguous(), R_last, t_last, R_delta, t_delta, huber_k, want_grad)
            return combine_icp(out, self.shard)

        def decode_set(self, owned: torch.Tensor) -> torch.Tensor:
            return expand_26(self._indexer, self._pos, self.n_xyz, owned)

        def owned_slots(self, slots: torch.Tensor) -> torch.Tensor:
            return slots[owner_of(self._pos[slots], self.shard.world) == self.shard.rank]

    return ShardedMap()
