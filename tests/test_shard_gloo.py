"""CPU, world_size 2, gloo: the multi-GPU host logic of difusion_b200/shard.py - super-block ownership and halo masks (against
brute force and against the C library's host mirror), the boundary-row exchange protocol (pack -> ONE all_to_all_single ->
unpack, restated with torch ops: the same buffers csrc/shard_xchg.cu fills on the GPU), the overflow value every rank must
agree on, and the ICP normal-equation combine.  The same ShardGroup code drives NCCL on GPUs."""
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def test_ownership_and_halo_masks_match_brute_force_and_the_library():
    from difusion_b200 import _lib, shard
    L = _lib.lib()
    n_xyz, k = [40, 24, 50], 3                                    # partial last blocks on every axis
    nx, ny, nz = n_xyz
    rng = np.random.default_rng(0)
    lin = np.unique(np.concatenate([rng.integers(0, nx * ny * nz, 4000), np.array([0, nx * ny * nz - 1, nz - 1, 7, 8])]))
    for world in (2, 8):
        own = shard.owner_of_np(lin, n_xyz, k, world)
        assert np.array_equal(own, shard.owner_of(torch.from_numpy(lin), n_xyz, k, world).numpy())
        assert np.array_equal(own[:200], np.array([L.dif_shard_owner(int(v), nx, ny, nz, k, world) for v in lin[:200]]))
        assert all((own == r).sum() > len(lin) // (3 * world) for r in range(world))            # balanced
        # halo mask == owners of the 27 cells around the cell (brute force over cells, not blocks)
        ix, iy, iz = lin // (ny * nz), (lin // nz) % ny, lin % nz
        brute = np.zeros(lin.shape, np.int64)
        for dx in (-1, 0, 1):
            for dy in (-1, 0, 1):
                for dz in (-1, 0, 1):
                    X, Y, Z = ix + dx, iy + dy, iz + dz
                    ok = (X >= 0) & (X < nx) & (Y >= 0) & (Y < ny) & (Z >= 0) & (Z < nz)
                    o = shard.owner_of_np(np.where(ok, (X * ny + Y) * nz + Z, 0), n_xyz, k, world)
                    brute |= np.where(ok, np.int64(1) << o, 0)
        assert np.array_equal(brute, shard.holder_mask_np(lin, n_xyz, k, world))
        assert np.all((brute >> own) & 1)                                                            # the owner always holds its row
    # interior cells of a block are held by their owner only
    inner = ((ix % 8 > 0) & (ix % 8 < 7) & (iy % 8 > 0) & (iy % 8 < 7) & (iz % 8 > 0) & (iz % 8 < 7))
    m8 = shard.holder_mask_np(lin, n_xyz, k, 8)
    assert np.array_equal(m8[inner], np.int64(1) << shard.owner_of_np(lin[inner], n_xyz, k, 8))


def _worker(rank, world, port, q):
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    from difusion_b200 import shard
    g = shard.ShardGroup()
    ok = True
    n_xyz, k = [48, 16, 48], 3
    nx, ny, nz = n_xyz
    # --- a "map": every occupied cell has a global slot (replicated integer state); rank r stores rows where bit r of the holder mask is set
    rng = np.random.default_rng(5)                                   # same stream on every rank = replicated state
    lin = np.sort(rng.choice(nx * ny * nz, 6000, replace=False))
    n_slots = lin.shape[0]
    own = shard.owner_of_np(lin, n_xyz, k, world)
    hold = shard.holder_mask_np(lin, n_xyz, k, world)
    stored = np.nonzero((hold >> rank) & 1)[0]
    row_of = torch.full((n_slots,), -1, dtype=torch.int64)
    row_of[torch.from_numpy(stored)] = torch.arange(len(stored))
    table = torch.zeros(len(stored), 32)
    truth = torch.arange(n_slots, dtype=torch.float32)[:, None] + torch.arange(29, dtype=torch.float32)[None, :] / 100
    ok &= len(stored) < 0.9 * n_slots                                  # the table really is a shard, not a replica

    def frame(updated_slots, cap):
        """owners fuse their updated rows, then one exchange; returns the overflow value."""
        mine = updated_slots[own[updated_slots] == rank]
        table[row_of[torch.from_numpy(mine)], :29] = truth[mine]
        masks = torch.from_numpy(hold[mine] & ~(1 << rank))
        send = shard.pack_reference(torch.from_numpy(mine), truth[mine], masks, rank, world, cap)
        recv = torch.zeros_like(send)
        g.all_to_all_fixed(send.view(-1), recv.view(-1))
        return shard.unpack_reference(table, row_of, recv, rank, cap)

    upd = np.sort(rng.choice(n_slots, 3000, replace=False))
    ov = frame(upd, cap=4096)
    ok &= ov == 0
    # every stored row that was updated equals the owner's value, halo rows included; nothing else was touched
    st_upd = np.intersect1d(stored, upd)
    ok &= bool(torch.equal(table[row_of[torch.from_numpy(st_upd)], :29], truth[st_upd]))
    untouched = np.setdiff1d(stored, upd)
    ok &= bool((table[row_of[torch.from_numpy(untouched)]] == 0).all())
    # --- overflow: a tiny buffer drops rows, every rank computes the SAME overflow value (the largest per-destination count of any sender)
    ov_small = frame(np.arange(n_slots), cap=8)
    both = torch.tensor([ov_small], dtype=torch.int64)
    gathered = [torch.zeros_like(both) for _ in range(world)]
    dist.all_gather(gathered, both)
    ok &= ov_small > 8 and all(int(t.item()) == ov_small for t in gathered)
    # recovery = republish with a large enough buffer: all stored rows become exact
    ov2 = frame(np.arange(n_slots), cap=ov_small)
    ok &= ov2 == 0 and bool(torch.equal(table[:, :29], truth[stored]))
    # --- ICP combine: per-rank normalised sums -> global normalised sums
    Ms = [3.0, 5.0]
    local = torch.zeros(44, dtype=torch.float64)
    local[:43] = (rank + 1) * torch.arange(43, dtype=torch.float64)
    local[43] = Ms[rank]
    comb = shard.combine_icp(local, g)
    expect = (Ms[0] * 1 * torch.arange(43, dtype=torch.float64) + Ms[1] * 2 * torch.arange(43, dtype=torch.float64)) / sum(Ms)
    ok &= bool(torch.allclose(comb[:43], expect)) and float(comb[43]) == sum(Ms)
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_sharded_exchange_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)], res
