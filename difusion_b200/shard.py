"""Hash-sharded PLIVox map over several GPUs (new design: the reference is single-GPU; SURVEY 8e, BASELINE configs[4]).

* Ownership: ``owner(cell) = splitmix64(super-block id) % world`` with super-blocks of ``2^k`` cells per axis (default 16^3 =
  80 cm cubes at 5 cm PLIVoxes).  Hashing blocks instead of cells keeps a PLIVox and (almost all of) its 26 neighbours on one rank.
* Integer state is REPLICATED: every rank receives the same frame and runs the same index kernels, so ``indexer``, slot numbering,
  ``latent_vecs_pos`` and ``voxel_obs_count`` are bit-identical to the single-GPU map on every rank (exact allocation needs the
  occupancy of non-owned neighbour cells).
* The floating-point payload is SHARDED: a rank stores latent rows only for the PLIVoxes it owns plus a one-cell halo around its
  super-blocks (what the marching-cubes blend of an owned PLIVox reads, mc_interp_kernel.cu:103-181), addressed through
  ``row_of_slot``; per-rank latent bytes are ~(1 + halo)/world of the map.  The encoder MLP and the fusion of a PLIVox run on its
  owner only.
* Per frame ONE exchange, ONE collective: ``dif_shard_pack`` puts the BOUNDARY rows this rank owns and fused (rows that sit on the
  surface of their super-block, i.e. in some other rank's halo) into per-destination segments of a fixed-size buffer with count
  headers; a single ``all_to_all_single`` with equal splits moves them; ``dif_shard_unpack`` writes the received rows into the
  halo.  No host synchronisation, sizes never leave the device.  A sender that ever has more rows for one destination than a
  segment holds is seen by EVERY rank (each receives every sender's header); the host notices a fixed number of frames later on
  all ranks at once, doubles the buffer and re-publishes all boundary rows.
* ICP linearisation: each rank linearises the points whose PLIVox it owns (``dif_shard_select_points`` compacts them), one
  all-reduce of 44 doubles.
* Mesh extraction: each rank meshes the PLIVoxes it owns; the cubes of their 26 neighbours are decoded locally from the halo.

``ShardGroup`` is pure ``torch.distributed`` plumbing (NCCL on GPUs; the same code runs on CPU tensors with gloo, which is how
tests/test_shard_gloo.py exercises the protocol).
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch
import torch.distributed as dist

_M64 = (1 << 64) - 1
XROW = 32


def _mix64_np(z):
    with np.errstate(over="ignore"):
        z = z.astype(np.uint64) + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def _n_blocks(n_xyz, k):
    return [((int(n) - 1) >> k) + 1 for n in n_xyz]


def block_owner_np(bx, by, bz, n_xyz, k: int, world: int) -> np.ndarray:
    """owner of super-block (bx, by, bz) - mirror of csrc/common.cuh shard_block_owner()."""
    _, nby, nbz = _n_blocks(n_xyz, k)
    bid = np.asarray(bz, np.int64) + nbz * (np.asarray(by, np.int64) + nby * np.asarray(bx, np.int64))
    return (_mix64_np(bid) % np.uint64(world)).astype(np.int64)


def owner_of_np(lin_ids, n_xyz, k: int, world: int) -> np.ndarray:
    """owner rank of the cells with the given linear ids (numpy) - mirror of shard_owner_lin()."""
    lin = np.asarray(lin_ids, np.int64)
    nx, ny, nz = [int(v) for v in n_xyz]
    return block_owner_np((lin // (ny * nz)) >> k, ((lin // nz) % ny) >> k, (lin % nz) >> k, n_xyz, k, world)


def owner_of(lin_ids: torch.Tensor, n_xyz, k: int, world: int) -> torch.Tensor:
    """The same function on a torch int64 tensor (any device): int64 arithmetic wraps, shifts are made logical by masking."""
    def lsr(v, s):
        return (v >> s) & ((1 << (64 - s)) - 1)

    def c(v):                     # python int -> wrapped int64 constant
        return v - (1 << 64) if v >= (1 << 63) else v
    nx, ny, nz = [int(v) for v in n_xyz]
    _, nby, nbz = _n_blocks(n_xyz, k)
    lin = lin_ids.to(torch.int64)
    bid = ((lin % nz) >> k) + nbz * ((((lin // nz) % ny) >> k) + nby * ((lin // (ny * nz)) >> k))
    z = bid + c(0x9E3779B97F4A7C15)
    z = (z ^ lsr(z, 30)) * c(0xBF58476D1CE4E5B9)
    z = (z ^ lsr(z, 27)) * c(0x94D049BB133111EB)
    z = z ^ lsr(z, 31)
    r = torch.remainder(z, world)                                       # unsigned modulo of a value stored in a signed int64
    return torch.where(z < 0, torch.remainder(r + ((1 << 64) % world), world), r)


def holder_mask_np(lin_ids, n_xyz, k: int, world: int) -> np.ndarray:
    """Bit r set: rank r keeps the cell's latent row (owner, or the cell is in r's halo) - mirror of shard_holder_mask()."""
    lin = np.asarray(lin_ids, np.int64)
    nx, ny, nz = [int(v) for v in n_xyz]
    ix, iy, iz = lin // (ny * nz), (lin // nz) % ny, lin % nz
    e = (1 << k) - 1
    m = np.zeros(lin.shape, np.int64)
    for dx in (-1, 0, 1):
        okx = (dx == 0) | ((dx < 0) & ((ix & e) == 0) & (ix > 0)) | ((dx > 0) & ((ix & e) == e) & (ix < nx - 1))
        for dy in (-1, 0, 1):
            oky = (dy == 0) | ((dy < 0) & ((iy & e) == 0) & (iy > 0)) | ((dy > 0) & ((iy & e) == e) & (iy < ny - 1))
            for dz in (-1, 0, 1):
                okz = (dz == 0) | ((dz < 0) & ((iz & e) == 0) & (iz > 0)) | ((dz > 0) & ((iz & e) == e) & (iz < nz - 1))
                ok = okx & oky & okz
                own = block_owner_np((ix >> k) + dx, (iy >> k) + dy, (iz >> k) + dz, n_xyz, k, world)
                m |= np.where(ok, np.int64(1) << own, 0)
    return m


class ShardGroup:
    def __init__(self, group=None):
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)

    def all_to_all_fixed(self, send: torch.Tensor, recv: torch.Tensor):
        """recv segment r = segment `rank` of rank r's send buffer; one collective, equal splits known statically."""
        dist.all_to_all_single(recv, send, group=self.group)

    def all_gather_rows(self, slots: torch.Tensor, rows: torch.Tensor):
        """Variable-length all-gather of (slot id, row) pairs (diagnostics / tests).  Returns the concatenation over ranks."""
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


# ---- the exchange protocol restated with torch ops (CPU / gloo test of the host logic; the GPU path is csrc/shard_xchg.cu) --------
def pack_reference(slots: torch.Tensor, rows29: torch.Tensor, dest_masks: torch.Tensor, rank: int, world: int, cap_rows: int) -> torch.Tensor:
    """[world][1 + cap][32] floats: what dif_shard_pack writes for the given (slot, row, holder mask) triples."""
    send = torch.zeros((world, cap_rows + 1, XROW), dtype=torch.float32)
    hdr = send.view(torch.int32)
    mx = 0
    for d in range(world):
        if d == rank:
            continue
        sel = torch.nonzero((dest_masks >> d) & 1).flatten()
        n = int(sel.numel())
        mx = max(mx, n)
        hdr[d, 0, 0] = n
        k = min(n, cap_rows)
        send.view(torch.int32)[d, 1:1 + k, 0] = slots[sel[:k]].to(torch.int32)
        send[d, 1:1 + k, 1:30] = rows29[sel[:k]]
    hdr[:, 0, 1] = mx
    return send


def unpack_reference(table: torch.Tensor, row_of: torch.Tensor, recv: torch.Tensor, rank: int, cap_rows: int) -> int:
    """Apply a received buffer to a local table; returns the overflow value dif_shard_unpack reports."""
    world = recv.size(0)
    hdr = recv.view(torch.int32)
    overflow = 0
    for src in range(world):
        count, smax = int(hdr[src, 0, 0]), int(hdr[src, 0, 1])
        if smax > cap_rows:
            overflow = max(overflow, smax)
        if src == rank:
            continue
        k = min(count, cap_rows)
        slots = hdr[src, 1:1 + k, 0].long()
        r = row_of[slots].long()
        ok = r >= 0
        table[r[ok], :29] = recv[src, 1:1 + k, 1:30][ok]
    return overflow


def expand_26(indexer: torch.Tensor, pos: torch.Tensor, n_xyz, slots: torch.Tensor) -> torch.Tensor:
    """Slots of the occupied cells in the 3x3x3 neighbourhood (own cell included) of the given slots, sorted unique.
    Owner-wise meshing decodes this set so that every cube the marching-cubes blend of an OWNED PLIVox can touch
    (mc_interp_kernel.cu:103-181: +-1 per axis, diagonals included) is in the decode batch, exactly as in a full single-GPU
    extraction; all of them are local rows (owned or halo)."""
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


def make_sharded_map(model, args, latent_dim, device, group: ShardGroup, block_log2: int = 4, **kw):
    from . import _lib
    from .system.map import DenseIndexedMap, LATENT_DIM

    class ShardedMap(DenseIndexedMap):
        RECOVERY_LAG = 4                                      # frames between raising the overflow flag and acting on it (all ranks alike)

        def __init__(self):
            super().__init__(model, args, latent_dim, device, shard=group, shard_block_log2=block_log2, **kw)
            self.shard = group
            self._xcap = 0
            self._xframe = 0
            self._xflag_ev = None
            self._sel_obs = None

        # ---------------------------------------------------------------- exchange buffers
        def _alloc_xchg(self, cap_rows: int):
            L = _lib.lib()
            self._xcap = cap_rows
            nfl = L.dif_shard_xchg_bytes(cap_rows, self.shard.world) // 4
            self._xsend = torch.zeros(nfl, dtype=torch.float32, device=self.device)
            self._xrecv = torch.zeros(nfl, dtype=torch.float32, device=self.device)
            self._xflag = torch.zeros(1, dtype=torch.int32, device=self.device)
            self._xflag_host = torch.zeros(1, dtype=torch.int32).pin_memory()
            self._xflag_ev = None

        def _exchange(self, n_x_dev: torch.Tensor):
            """pack -> ONE all-to-all -> unpack, all on the current stream."""
            L = _lib.lib()
            view, st = self._view(), _lib.stream_ptr(self.device)
            _lib.check(L.dif_shard_pack(ctypes.byref(view), n_x_dev.data_ptr(), self._xcap, self._xsend.data_ptr(), st), "dif_shard_pack")
            self.shard.all_to_all_fixed(self._xsend, self._xrecv)
            _lib.check(L.dif_shard_unpack(ctypes.byref(view), self._xrecv.data_ptr(), self._xcap, self._xflag.data_ptr(), st), "dif_shard_unpack")

        def resync(self):
            """Publish ALL boundary rows this rank owns (recovery after an exchange-buffer overflow); grows the buffer until they fit."""
            n_occ = self.n_occupied
            owned = self.owned_slots(torch.arange(n_occ, device=self.device)).to(torch.int32)
            while True:
                need = owned.numel() + 1
                if self._xchg.numel() < need:
                    self._xchg = torch.empty(max(need, 2 * self._xchg.numel()), dtype=torch.int32, device=self.device)
                self._xchg[:owned.numel()] = owned
                n_x = torch.tensor([owned.numel()], dtype=torch.int32, device=self.device)
                self._xflag.zero_()
                self._exchange(n_x)
                worst = int(self._xflag.item())                       # identical on every rank: the loop is collective
                if worst <= self._xcap:
                    break
                cap = self._xcap
                while cap < worst:
                    cap *= 2
                self._alloc_xchg(cap)

        # ---------------------------------------------------------------- integrate (map.py:340-519) + exchange
        def integrate_keyframe(self, surface_xyz, surface_normal, do_optimize=False, async_optimize=False):
            need = min(8 * surface_xyz.size(0), self._cap_phys) + 1
            if self._xchg.numel() < need:
                self._xchg = torch.empty(need, dtype=torch.int32, device=self.device)
                self._view_key = None
            if self._xcap == 0:
                self._alloc_xchg(1 << 13)
            self._xframe += 1
            if self._xflag_ev is not None and self._xframe >= self._xflag_frame + self.RECOVERY_LAG:
                # did an exchange overflow?  Checked a fixed number of frames after the flag was copied (long complete, so the
                # wait is free) so that every rank takes the collective recovery path in the same frame; the flag itself is
                # identical on every rank because every rank sees every sender's header.
                self._xflag_ev.synchronize()
                self._xflag_ev = None
                worst = int(self._xflag_host[0])                      # largest per-destination row count any rank tried to send
                if worst:
                    cap = self._xcap * 2
                    while cap < 2 * worst:
                        cap *= 2
                    self._alloc_xchg(cap)
                    self.resync()
            mask = super().integrate_keyframe(surface_xyz, surface_normal, do_optimize, async_optimize)
            self._exchange(self._stats_dev[_lib.STAT_N_XCHG:])
            if self._xflag_ev is None:
                self._xflag_host.copy_(self._xflag, non_blocking=True)
                self._xflag_ev = torch.cuda.Event()
                self._xflag_ev.record(torch.cuda.current_stream(self.device))
                self._xflag_frame = self._xframe
            return mask

        @property
        def last_exchange(self):
            """Rows this rank sent / received in the last frame (reads the headers: host sync; diagnostics only)."""
            w = self.shard.world
            sent = self._xsend.view(w, -1)[:, 0].contiguous().view(torch.int32).tolist()
            got = self._xrecv.view(w, -1)[:, 0].contiguous().view(torch.int32).tolist()
            return dict(rows_sent=sum(sent), rows_received=sum(got), per_destination=sent, capacity=self._xcap,
                        bytes_per_rank=self._xsend.numel() * 4)

        # ---------------------------------------------------------------- storage views
        @property
        def n_rows(self) -> int:
            """latent rows stored on this rank (owned + halo)."""
            return int(self._n_rows_dev.item())

        def local_latents(self):
            """(slots, rows (k, 29)) of every PLIVox whose latent row lives on this rank."""
            n = self.n_occupied
            r = self._row_of[:n].long()
            s = torch.nonzero(r >= 0).flatten()
            return s, self._latent[r[s], :LATENT_DIM]

        @property
        def latent_vecs(self):
            """The reference's (capacity, 29) tensor restricted to what this rank stores: rows of non-local PLIVoxes are zero."""
            cap = self._cap_ref()
            out = torch.zeros((cap, LATENT_DIM), dtype=torch.float32, device=self.device)
            s, rows = self.local_latents()
            out[s] = rows
            return out

        def gather_latents(self):
            """The full (capacity, 29) table assembled from every rank's OWNED rows (collective; tests / save)."""
            owned = self.owned_slots(torch.arange(self.n_occupied, device=self.device))
            rows = self._latent[self._row_of[owned].long(), :LATENT_DIM]
            s_all, r_all = self.shard.all_gather_rows(owned.int(), rows.contiguous())
            out = torch.zeros((self._cap_ref(), LATENT_DIM), dtype=torch.float32, device=self.device)
            out[s_all.long()] = r_all
            return out

        def owned_slots(self, slots: torch.Tensor) -> torch.Tensor:
            return slots[owner_of(self._pos[slots], self.n_xyz, self._shard_k, self.shard.world) == self.shard.rank]

        def decode_set(self, owned: torch.Tensor) -> torch.Tensor:
            return expand_26(self._indexer, self._pos, self.n_xyz, owned)

        def _snapshot(self):
            snap = super()._snapshot()
            snap["tensors"] = snap["tensors"] + (self._row_of, self._n_rows_dev)
            return snap

        # ---------------------------------------------------------------- tracker term: owned points only, then one all-reduce
        def icp_linearize(self, obs_xyz, R_last, t_last, R_delta, t_delta, huber_k=5.0, want_grad=True):
            L = _lib.lib()
            x = obs_xyz.detach().contiguous().float()
            n = x.size(0)
            if self._sel_obs is None or self._sel_obs.size(0) < n:
                self._sel_obs = torch.empty((max(n, 1 << 15), 3), dtype=torch.float32, device=self.device)
                self._sel_frame = torch.zeros(_lib.FRAME_HEADER_FLOATS, dtype=torch.float32, device=self.device)
            if self._icp_scratch is None:
                self._icp_scratch = torch.zeros(L.dif_icp_scratch_bytes(n), dtype=torch.uint8, device=self.device)
                self._icp_ring = torch.empty((16, 44), dtype=torch.float64, device=self.device)
                self._icp_next = 0
            out = self._icp_ring[self._icp_next]
            self._icp_next = (self._icp_next + 1) % 16
            pose = np.empty(24, np.float32)
            pose[0:9], pose[9:12], pose[12:21], pose[21:24] = np.ravel(R_last), np.ravel(t_last), np.ravel(R_delta), np.ravel(t_delta)
            view, st = self._view(), _lib.stream_ptr(self.device)
            _lib.check(L.dif_shard_select_points(ctypes.byref(view), x.data_ptr(), n, pose.ctypes.data, self._sel_obs.data_ptr(),
                                                 self._sel_frame.data_ptr(), st), "dif_shard_select_points")
            _lib.check(L.dif_icp_linearize(ctypes.byref(view), self._prep.decoder.data_ptr(), self._sel_obs.data_ptr(), n, None,
                                           self._sel_frame.data_ptr(), float(huber_k) if huber_k else 0.0, int(want_grad),
                                           self._icp_scratch.data_ptr(), self._icp_scratch.numel(), out.data_ptr(), st), "dif_icp_linearize")
            return combine_icp(out, self.shard)

        def get_sdf(self, xyz: torch.Tensor):
            """map.py:559-579 on the sharded map: every rank decodes the points whose PLIVox it owns; two all-reduces assemble
            the single-GPU answer (no autograd through the collective: the sharded tracker uses icp_linearize)."""
            slot, rel = self._query(xyz)                                     # local ROW of the owner, -1 elsewhere
            L = _lib.lib()
            n = xyz.size(0)
            valid = (slot >= 0).to(torch.float32)
            sdf = torch.zeros(n, dtype=torch.float32, device=self.device)
            std = torch.zeros(n, dtype=torch.float32, device=self.device)
            _lib.check(L.dif_decode(self._prep.decoder.data_ptr(), self._latent.data_ptr(), _lib.LATENT_ROW_FLOATS, slot.data_ptr(),
                                    rel.data_ptr(), n, None, 1.0, sdf.data_ptr(), std.data_ptr(), None, None, _lib.stream_ptr(self.device)),
                       "dif_decode")
            packed = torch.stack([sdf, std, valid])
            self.shard.all_reduce_sum(packed)
            v = packed[2] > 0
            assert bool(v.any())
            return packed[0][v], packed[1][v], v

    return ShardedMap()
