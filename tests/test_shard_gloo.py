"""CPU, world_size 2, gloo: the multi-GPU host logic of difusion_b200/shard.py (ownership hash, variable-length row exchange,
ICP normal-equation combine) - the same code drives NCCL on GPUs."""
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, q):
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    from difusion_b200 import shard
    g = shard.ShardGroup()
    ok = True
    # --- variable-length exchange, including an empty contribution
    rng = np.random.default_rng(100 + rank)
    for counts in ([5, 9], [0, 4], [0, 0], [300, 1]):
        n = counts[rank]
        slots = torch.from_numpy(rng.integers(0, 1000, n).astype(np.int32))
        rows = torch.from_numpy(rng.normal(size=(n, 29)).astype(np.float32))
        s_all, r_all = g.all_gather_rows(slots, rows)
        ok &= s_all.numel() == sum(counts) and r_all.shape == (sum(counts), 29)
        off = sum(counts[:rank])
        ok &= torch.equal(s_all[off:off + n], slots) and torch.equal(r_all[off:off + n], rows)
    # --- a sharded "map": every rank fuses the rows it owns, the exchange makes all replicas identical
    world_rows = 4096
    lin = torch.arange(world_rows, dtype=torch.int64) * 7 + 3
    own = shard.owner_of(lin, world)
    ok &= bool(torch.equal(own, torch.from_numpy(shard.owner_of_np(lin.numpy(), world))))
    ok &= all(int((own == r).sum()) > world_rows // (2 * world) for r in range(world))       # balanced
    table = torch.zeros(world_rows, 29)
    mine = torch.nonzero(own == rank).flatten()
    table[mine] = torch.arange(world_rows, dtype=torch.float32)[mine, None] + torch.arange(29, dtype=torch.float32)[None, :] / 100
    s_all, r_all = g.all_gather_rows(mine.int(), table[mine])
    table.index_copy_(0, s_all.long(), r_all)
    ref = torch.arange(world_rows, dtype=torch.float32)[:, None] + torch.arange(29, dtype=torch.float32)[None, :] / 100
    ok &= bool(torch.equal(table, ref))
    # --- the per-frame exchange protocol (dif_shard_pack / one all_gather_into_tensor / dif_shard_unpack), restated with torch ops:
    #     [1 + cap][32] floats per rank, row 0 word 0 = row count (int bits), row 1+i = slot bits, 29 latents, 2 pad words
    cap = 64
    for counts in ([10, 64], [0, 3], [70, 5]):                                   # the last one overflows rank 0's buffer
        n = counts[rank]
        slots = torch.from_numpy(rng.choice(1000, n, replace=False).astype(np.int32)) + 1000 * rank
        rows = torch.from_numpy(rng.normal(size=(n, 29)).astype(np.float32))
        send = torch.zeros((cap + 1) * 32)
        send[0] = torch.tensor([n], dtype=torch.int32).view(torch.float32)[0]
        k = min(n, cap)
        body = send[32:].view(cap, 32)
        body[:k, 0] = slots[:k].view(torch.float32)
        body[:k, 1:30] = rows[:k]
        recv = torch.zeros(world * (cap + 1) * 32)
        g.all_gather_fixed(send, recv)
        table, overflow = torch.zeros(2000, 29), 0
        for src in range(world):
            buf = recv.view(world, cap + 1, 32)[src]
            cnt = int(buf[0, :1].view(torch.int32)[0])
            overflow = max(overflow, cnt if cnt > cap else 0)
            if src != rank:
                kk = min(cnt, cap)
                table[buf[1:1 + kk, 0].contiguous().view(torch.int32).long()] = buf[1:1 + kk, 1:30]
        ok &= overflow == (70 if max(counts) > cap else 0)                       # every rank sees the same overflow value
        s_all, r_all = g.all_gather_rows(slots[:k], rows[:k])                    # what the other rank really published (first cap rows)
        other = 1 - rank
        off, ko = (0 if other == 0 else min(counts[0], cap)), min(counts[other], cap)
        ok &= bool(torch.equal(table[s_all[off:off + ko].long()], r_all[off:off + ko]))
        ok &= int((table.abs().sum(1) > 0).sum()) == ko
    # --- ICP combine: per-rank (already normalised) partial systems -> the global one
    rng = np.random.default_rng(7)
    J = rng.normal(size=(1000, 6)); r = rng.normal(size=1000)
    lo, hi = 1000 * rank // world, 1000 * (rank + 1) // world
    Jr, rr = J[lo:hi], r[lo:hi]
    out = np.zeros(44); M = hi - lo
    out[:36] = (Jr.T @ Jr / M).ravel(); out[36:42] = Jr.T @ rr / M; out[42] = rr @ rr / M; out[43] = M
    tot = shard.combine_icp(torch.from_numpy(out), g).numpy()
    ok &= np.allclose(tot[:36], (J.T @ J / 1000).ravel()) and np.allclose(tot[36:42], J.T @ r / 1000) and np.isclose(tot[42], r @ r / 1000) and tot[43] == 1000
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_shard_host_logic_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


def test_expand_26_matches_brute_force():
    """Owner-wise meshing decodes the owned PLIVoxes plus every occupied cell in their 3x3x3 neighbourhoods (shard.expand_26)."""
    from difusion_b200 import shard
    rng = np.random.default_rng(4)
    n_xyz = [7, 5, 6]
    n_cells = int(np.prod(n_xyz))
    occ = np.sort(rng.choice(n_cells, 90, replace=False))
    order = rng.permutation(occ.size)                                    # slot numbering is arbitrary
    indexer = np.full(n_cells, -1, np.int64)
    pos = np.full(128, -1, np.int64)
    indexer[occ[order]] = np.arange(occ.size)
    pos[:occ.size] = occ[order]
    owned = np.sort(rng.choice(occ.size, 25, replace=False))
    got = shard.expand_26(torch.from_numpy(indexer), torch.from_numpy(pos), n_xyz, torch.from_numpy(owned)).numpy()
    exp = set()
    for s in owned:
        lin = pos[s]
        x, y, z = lin // (n_xyz[1] * n_xyz[2]), (lin // n_xyz[2]) % n_xyz[1], lin % n_xyz[2]
        for dx in (-1, 0, 1):
            for dy in (-1, 0, 1):
                for dz in (-1, 0, 1):
                    X, Y, Z = x + dx, y + dy, z + dz
                    if 0 <= X < n_xyz[0] and 0 <= Y < n_xyz[1] and 0 <= Z < n_xyz[2]:
                        t = indexer[(X * n_xyz[1] + Y) * n_xyz[2] + Z]
                        if t >= 0:
                            exp.add(int(t))
    assert got.tolist() == sorted(exp) and set(owned.tolist()) <= exp
    assert shard.expand_26(torch.from_numpy(indexer), torch.from_numpy(pos), n_xyz, torch.zeros(0, dtype=torch.int64)).numel() == 0
