"""ORACLE - CPU restatement of the DI-Fusion per-frame hot path.  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs
may import this module; nothing under ``difusion_b200/`` does (the product path fails loudly without
its CUDA library instead of falling back here).

What it restates (reference = huangjh-pub/di-fusion @ dd6ab8e, paths relative to ``pytorch/``):

=====================  ==========================================  =================================
function here          reference                                   SURVEY row
=====================  ==========================================  =================================
fold_decoder           network/di_decoder.py:10-47 + weight_norm   a-7
fold_encoder           network/di_encoder.py:6-12, pt_util:37-116  a-5
decoder_forward        network/di_decoder.py:55-86                 a-7
encoder_forward        network/di_encoder.py:26-30                 a-5
OracleMap.integrate    system/map.py:340-519                       a-2 .. a-6
OracleMap.get_sdf      system/map.py:559-579                       a-8
OracleMap.mesh_cubes   system/map.py:624-687, utility.py:129-149   a-10
OracleMap.extract      system/map.py:689-702 + oracle/mc_oracle.c  a-11
compute_sdf_Hg         system/tracker.py:174-218                   a-9
=====================  ==========================================  =================================

Arithmetic that lives in a third-party dependency of the reference (PyTorch, unpinned in
``requirements.txt:1``; this image has torch 2.11.0): Linear/Conv1d/BatchNorm/unique/interpolate/autograd.
The restatement calls the same torch CPU fp32 operators for the floating-point MLP and trilinear parts
(so it times like the reference on CPU) and numpy for all integer / index work.

PINNING: tests/golden/make_golden.py runs the UNMODIFIED reference (oracle/ref_shim.py, build container
only) and this restatement on the same seeded inputs; tests/test_oracle_golden.py re-checks this file
against the committed outputs of that run on every machine.  Integer state is compared bit-exactly.
The marching-cubes stage has no executable reference on CPU: "parity unpinned" for that stage only
(see oracle/mc_oracle.c).
"""
from __future__ import annotations

import math
from types import SimpleNamespace

import numpy as np
import torch
import torch.nn.functional as F

LATENT_DIM = 29


# ----------------------------------------------------------------------------------------- weights
def fold_decoder(sd: dict) -> SimpleNamespace:
    """Weight-norm folding W = g * v / ||v||_row  (torch.nn.utils.weight_norm, dim=0), di_decoder.py:36-40."""
    out = SimpleNamespace(W=[], b=[])
    for k in range(5):
        v = sd[f"lin{k}.weight_v"].float()
        g = sd[f"lin{k}.weight_g"].float()
        W = v * (g / v.norm(dim=1, keepdim=True))
        out.W.append(W.contiguous())
        out.b.append(sd[f"lin{k}.bias"].float().contiguous())
    out.Wu = sd["uncertainty_layer.weight"].float().contiguous()      # (1,128), plain Linear (di_decoder.py:47)
    out.bu = sd["uncertainty_layer.bias"].float().contiguous()
    return out


def fold_encoder(sd: dict, eps: float = 1e-5) -> SimpleNamespace:
    """Conv1d(k=1, no bias) + BatchNorm1d(eval) folded to W' x + b'  (pt_util.py:37-43,83-116,193-206)."""
    out = SimpleNamespace(W=[], b=[])
    for k in range(3):
        W = sd[f"mlp.layer{k}.conv.weight"].float().squeeze(-1)
        bn = f"mlp.layer{k}.normlayer.bn."
        scale = sd[bn + "weight"].float() / torch.sqrt(sd[bn + "running_var"].float() + eps)
        out.W.append((W * scale[:, None]).contiguous())
        out.b.append((sd[bn + "bias"].float() - sd[bn + "running_mean"].float() * scale).contiguous())
    out.W.append(sd["mlp.layer3.conv.weight"].float().squeeze(-1).contiguous())
    out.b.append(sd["mlp.layer3.conv.bias"].float().contiguous())
    return out


def load_weights_npz(path) -> SimpleNamespace:
    """Raw (unfolded) checkpoint tensors exported by tests/golden/make_golden.py -> folded weights."""
    z = np.load(path)
    dec = {k[4:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("dec.")}
    enc = {k[4:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("enc.")}
    return SimpleNamespace(dec=fold_decoder(dec), enc=fold_encoder(enc), raw_dec=dec, raw_enc=enc)


# ----------------------------------------------------------------------------------------- networks
def decoder_forward(dec, latent: torch.Tensor, xyz: torch.Tensor):
    """di_decoder.py:55-86 in eval mode.  latent (n,29), xyz (n,3) -> sdf (n,), std (n,)."""
    x_in = torch.cat([latent, xyz], dim=1)                 # utility.py:80
    h = F.relu(F.linear(x_in, dec.W[0], dec.b[0]))
    h = F.relu(F.linear(h, dec.W[1], dec.b[1]))
    h = F.relu(F.linear(h, dec.W[2], dec.b[2]))
    h = F.relu(F.linear(torch.cat([h, x_in], 1), dec.W[3], dec.b[3]))     # latent_in = [3], h first (:61-62)
    std = 0.05 + 0.5 * F.softplus(F.linear(h, dec.Wu, dec.bu))             # :65-68
    sdf = torch.tanh(F.linear(h, dec.W[4], dec.b[4]))                      # :70,84
    return sdf.squeeze(-1), std.squeeze(-1)


def encoder_forward(enc, xyzn: torch.Tensor):
    """di_encoder.py:26-30 ('cnp' mode, eval).  (S,6) -> (S,29)."""
    h = xyzn
    for k in range(3):
        h = F.relu(F.linear(h, enc.W[k], enc.b[k]))
    return F.linear(h, enc.W[3], enc.b[3])


def get_samples(r: int, a: float = 0.0, b: float | None = None) -> np.ndarray:
    """utility.py:129-149: x-major r^3 lattice over [a, b], fp32."""
    idx = torch.arange(0, r ** 3, dtype=torch.long)
    if b is None:
        b = 1.0 - 1.0 / r
    vsize = (b - a) / (r - 1)
    s = torch.zeros(r ** 3, 3, dtype=torch.float32)
    s[:, 0] = (idx // (r * r)) * vsize + a
    s[:, 1] = ((idx // r) % r) * vsize + a
    s[:, 2] = (idx % r) * vsize + a
    return s


# ----------------------------------------------------------------------------------------- the map
_OFFS = np.array([[-.5, -.5, -.5], [-.5, -.5, .5], [-.5, .5, -.5], [-.5, .5, .5],
                  [.5, -.5, -.5], [.5, -.5, .5], [.5, .5, -.5], [.5, .5, .5]], np.float32)     # map.py:186-189
_NBR6 = np.array([[-1, 0, 0], [1, 0, 0], [0, -1, 0], [0, 1, 0], [0, 0, -1], [0, 0, 1]], np.int64)   # map.py:193-196


class OracleMap:
    def __init__(self, weights, args):
        self.w = weights
        self.args = args
        self.voxel_size = args.voxel_size
        # map.py:178
        self.n_xyz = np.ceil((np.asarray(args.bound_max) - np.asarray(args.bound_min)) / args.voxel_size).astype(int).tolist()
        self.bound_min = np.asarray(args.bound_min, dtype=np.float32)
        self.n_cells = int(np.prod(self.n_xyz))
        self.n_occupied = 0
        self.indexer = np.full(self.n_cells, -1, np.int64)                 # map.py:201
        self.latent_vecs = np.zeros((1, LATENT_DIM), np.float32)           # map.py:204,234
        self.latent_vecs_pos = np.full((1,), -1, np.int64)
        self.voxel_obs_count = np.zeros((1,), np.float32)
        self.voxel_optimized = np.zeros((1,), bool)
        self.updated_vec_id = np.zeros((0,), np.int64)                     # mesh_cache.updated_vec_id
        self.last_stats = {}

    # -- addressing ------------------------------------------------------------------------
    def _lin(self, ijk):                                                   # map.py:287-292
        return ijk[:, 2] + self.n_xyz[2] * ijk[:, 1] + (self.n_xyz[2] * self.n_xyz[1]) * ijk[:, 0]

    def _unlin(self, idx):                                                 # map.py:294-301
        return np.stack([idx // (self.n_xyz[1] * self.n_xyz[2]), (idx // self.n_xyz[2]) % self.n_xyz[1],
                         idx % self.n_xyz[2]], axis=-1)

    def _normalize(self, xyz: np.ndarray) -> np.ndarray:                   # map.py:366-367 (fp32 sub, fp32 true div)
        return ((xyz.astype(np.float32) - self.bound_min[None, :]) / np.float32(self.voxel_size)).astype(np.float32)

    def _dilate6(self, ids: np.ndarray, only_allocated: bool) -> np.ndarray:   # map.py:545-557
        parts = [ids]
        ijk = self._unlin(ids)
        hi = np.asarray(self.n_xyz, np.int64) - 1
        for d in _NBR6:
            n = self._lin(np.clip(ijk + d[None, :], 0, hi[None, :]))
            if only_allocated:
                n = n[self.indexer[n] != -1]
            parts.append(n)
        return np.unique(np.concatenate(parts))

    def _grow(self, count: int) -> np.ndarray:                             # map.py:263-285
        target = self.n_occupied + count
        cap = self.latent_vecs.shape[0]
        if cap < target:
            new = cap
            while new < target:
                new *= 2
            lv = np.zeros((new, LATENT_DIM), np.float32); lv[:cap] = self.latent_vecs
            pos = np.full((new,), -1, np.int64); pos[:cap] = self.latent_vecs_pos
            cnt = np.zeros((new,), np.float32); cnt[:cap] = self.voxel_obs_count
            opt = np.zeros((new,), bool); opt[:cap] = self.voxel_optimized
            self.latent_vecs, self.latent_vecs_pos, self.voxel_obs_count, self.voxel_optimized = lv, pos, cnt, opt
        ids = np.arange(self.n_occupied, target, dtype=np.int64)
        self.n_occupied = target
        return ids

    def allocate_block(self, lin_ids: np.ndarray):                         # map.py:310-319
        slots = self._grow(lin_ids.shape[0])
        self.latent_vecs_pos[slots] = lin_ids
        self.indexer[lin_ids] = slots

    # -- a-2 .. a-6 -------------------------------------------------------------------------
    def integrate_keyframe(self, surface_xyz: np.ndarray, surface_normal: np.ndarray):
        a = self.args
        p = self._normalize(surface_xyz)
        cell = self._lin(np.ceil(p).astype(np.int64) - 1)                  # map.py:368-369
        normal = surface_normal.astype(np.float32)

        unq_mask = None
        if a.prune_min_vox_obs > 0:                                        # map.py:373-378
            _, inv, cnt = np.unique(cell, return_inverse=True, return_counts=True)
            unq_mask = (cnt > a.prune_min_vox_obs)[inv]
            p, cell, normal = p[unq_mask], cell[unq_mask], normal[unq_mask]

        empty = self.indexer[cell] == -1                                   # map.py:381-387
        n_new = 0
        if empty.sum() > 0:
            new_ids = self._dilate6(np.unique(cell[empty]), only_allocated=False)
            new_ids = new_ids[self.indexer[new_ids] == -1]
            n_new = new_ids.shape[0]
            self.allocate_block(new_ids)

        # map.py:407-411: encoder targets T
        enc_pos = self.latent_vecs_pos[np.logical_and(self.voxel_obs_count < a.encoder_count_th, self.latent_vecs_pos >= 0)]
        status = np.zeros(self.n_cells, np.int16)
        status[enc_pos] |= 1
        self.last_stats = dict(n_kept=int(p.shape[0]), n_new=int(n_new), n_samples=0, n_updated=0)
        if enc_pos.shape[0] > 0:
            focus = np.zeros(self.n_cells, np.int64)                       # map.py:389-397
            focus[self._dilate6(enc_pos, only_allocated=False)] = 1
            fm = focus[cell] == 1
            p_f, n_f = p[fm], normal[fm]

            hi = (np.asarray(self.n_xyz, np.float32) - 1)[None, :]
            g_slot, g_xyzn = [], []
            for off in _OFFS:                                              # map.py:421-433
                c = np.clip(np.ceil(p_f + off[None, :]) - np.float32(1), np.float32(0), hi)
                rel = (p_f - c) - np.float32(0.5)
                lin = self._lin(c.astype(np.int64))
                m = status[lin] >= 1
                g_slot.append(self.indexer[lin][m])
                g_xyzn.append(np.concatenate([rel[m], n_f[m]], axis=-1))
            g_slot = np.concatenate(g_slot)
            g_xyzn = np.concatenate(g_xyzn).astype(np.float32)
            uniq, inv, cnt = np.unique(g_slot, return_inverse=True, return_counts=True)   # map.py:437-439
            self.last_stats.update(n_samples=int(g_slot.shape[0]), n_updated=int(uniq.shape[0]))
            if g_slot.shape[0] > 0:
                with torch.no_grad():
                    enc = encoder_forward(self.w.enc, torch.from_numpy(g_xyzn))           # map.py:446
                    s = torch.zeros(uniq.shape[0], LATENT_DIM).index_add_(0, torch.from_numpy(inv), enc)   # :448
                    s = s.numpy()
                pc = cnt.astype(np.float32)
                s = s + self.latent_vecs[uniq] * self.voxel_obs_count[uniq][:, None]      # map.py:449
                self.voxel_obs_count[uniq] += pc                                          # :450
                self.latent_vecs[uniq] = s / self.voxel_obs_count[uniq][:, None]          # :451
                self.updated_vec_id = np.unique(np.concatenate([self.updated_vec_id, uniq]))   # :303-308
        return unq_mask

    # -- a-8 ----------------------------------------------------------------------------------
    def lookup(self, xyz: np.ndarray):
        """map.py:565-575: (slot per valid point, rel xyz of valid points, valid mask)."""
        p = self._normalize(xyz)
        gid = np.ceil(p).astype(np.int64) - 1
        slot = self.indexer[self._lin(gid)]
        valid = slot != -1
        vv = self.voxel_obs_count[slot[valid]] > self.args.ignore_count_th
        valid[valid.copy()] = vv
        rel = (p[valid] - gid[valid].astype(np.float32)) - np.float32(0.5)
        return slot[valid], rel.astype(np.float32), valid

    def get_sdf(self, xyz: np.ndarray, want_grad: bool = False):
        """-> sdf (M,), std (M,), valid (N,) [, d(sdf/std.detach())/d xyz_world (M,3)]  (tracker.py:186-194)."""
        slot, rel, valid = self.lookup(xyz)
        assert slot.shape[0] > 0, "reference asserts on an empty batch (utility.py:84-85)"
        lat = torch.from_numpy(self.latent_vecs[slot])
        x = torch.from_numpy(rel).requires_grad_(want_grad)
        sdf, std = decoder_forward(self.w.dec, lat, x)
        if not want_grad:
            return sdf.detach().numpy(), std.detach().numpy(), valid
        r = sdf / std.detach()
        g = torch.autograd.grad(r, [x], grad_outputs=torch.ones_like(r))[0]
        g = g / np.float32(self.voxel_size)              # chain rule through xyz_normalized (map.py:565)
        return sdf.detach().numpy(), std.detach().numpy(), valid, g.numpy()

    # -- a-10 ---------------------------------------------------------------------------------
    def mesh_cubes(self, voxel_resolution: int, fast: bool = True, no_cache: bool = True):
        """map.py:614-687 -> (focused_flatten_id, vec_id_batch_mapping, high_sdf(negated), high_std, occupied_vec_id)."""
        if no_cache:
            updated = np.arange(self.n_occupied, dtype=np.int64)
        else:
            updated = self.updated_vec_id
            self.updated_vec_id = np.zeros((0,), np.int64)
        focused = self.latent_vecs_pos[updated]
        occ = self.indexer[self._dilate6(focused, only_allocated=True)]
        occ = occ[self.voxel_obs_count[occ] > self.args.ignore_count_th]
        mapping = np.full((int(occ.max()) + 1,), -1, np.int32)
        mapping[occ] = np.arange(occ.shape[0], dtype=np.int32)
        lat = torch.from_numpy(self.latent_vecs[occ])
        B = lat.shape[0]
        r = voxel_resolution
        sa = -(r // 2) * (1. / r)
        sb = 1. + (r - 1) // 2 * (1. / r)
        hr = 2 * r
        lr = hr // 2 if fast else hr
        with torch.no_grad():
            ls = get_samples(lr, sa, sb) - 0.5
            low_sdf, low_std = decoder_forward(self.w.dec, lat.unsqueeze(1).repeat(1, lr ** 3, 1).view(-1, LATENT_DIM),
                                               ls.unsqueeze(0).repeat(B, 1, 1).view(-1, 3))
            if fast:
                low_sdf = low_sdf.reshape(B, 1, lr, lr, lr)
                low_std = low_std.reshape(B, 1, lr, lr, lr)
                hs = F.interpolate(low_sdf, mode="trilinear", size=(hr, hr, hr), align_corners=True).reshape(B, hr ** 3)
                hd = F.interpolate(low_std, mode="trilinear", size=(hr, hr, hr), align_corners=True).reshape(B, hr ** 3)
                li, si = torch.where(hs.abs() < 0.05)
                n_high = int(li.shape[0])
                if n_high > 0:
                    hsmp = get_samples(hr, sa, sb) - 0.5
                    v_sdf, v_std = decoder_forward(self.w.dec, lat[li], hsmp[si])
                    hs[li, si] = v_sdf
                    hd[li, si] = v_std
                hs = hs.reshape(B, hr, hr, hr)
                hd = hd.reshape(B, hr, hr, hr)
            else:
                n_high = 0
                hs = low_sdf.reshape(B, lr, lr, lr)
                hd = low_std.reshape(B, lr, lr, lr)
            hs = -hs                                                       # map.py:687
        self.last_stats = dict(B=B, n_low=B * lr ** 3, n_high=n_high)
        return focused, mapping, hs.numpy(), hd.numpy(), occ

    # -- a-11 ---------------------------------------------------------------------------------
    def extract_mesh(self, voxel_resolution: int, max_n_triangles: int, fast: bool = True, max_std: float = 2000.0,
                     no_cache: bool = True):
        """map.py:689-702: triangles in world units (T,3,3), flatten id (T,), std (T,3)."""
        from oracle import mc_oracle
        focused, mapping, hs, hd, _ = self.mesh_cubes(voxel_resolution, fast, no_cache)
        tri, fid, std = mc_oracle.marching_cubes_interp(self.indexer.reshape(self.n_xyz), focused, mapping, hs, hd,
                                                        max_n_triangles, self.n_xyz, max_std)
        tri = tri * np.float32(self.voxel_size) + self.bound_min[None, None, :]
        return tri, fid, std


# ----------------------------------------------------------------------------------------- a-12 / f-2
def host_cache_keep_mask(cache_ids: np.ndarray, new_ids: np.ndarray) -> np.ndarray:
    """map.py:708-709 + _get_valid_idx (:20-26): a cached triangle survives iff its PLIVox id is not among the new ones.
    (The reference searches the sorted unique new ids; for a cached id above their maximum it reads one past the end of the
    array - undefined under numba - where the intent, and this restatement, is "keep".)  Pinned against the executed reference
    by tests/golden/make_golden_merge.py -> tests/golden/ref_host_merge.npz."""
    return ~np.isin(cache_ids, np.unique(new_ids))


def host_cache_merge(cache, new, voxel_size: np.float32, bound_min: np.ndarray):
    """map.py:698-714: voxel units -> world (two rounded fp32 ops), drop cached rows of re-meshed PLIVoxes, append.
    cache / new: (tri (T,3,3) f32, id (T,) i64, std (T,3) f32); cache may be None."""
    world = (new[0] * np.float32(voxel_size) + bound_min.astype(np.float32), new[1], new[2])
    if cache is None:
        return world
    keep = host_cache_keep_mask(cache[1], new[1])
    return tuple(np.concatenate([c[keep], n], 0) for c, n in zip(cache, world))


# ----------------------------------------------------------------------------------------- a-9
def compose(Ra, ta, Rb, tb):
    """Isometry.dot (motion_util.py:277-278) on rotation matrices, float64."""
    return Ra @ Rb, Ra @ tb + ta


def huber_weight(x: torch.Tensor, k: float) -> torch.Tensor:             # tracker.py:59-65
    w = torch.ones_like(x)
    ax = x.abs()
    m = ax > k
    w[m] = k / ax[m]
    return w


def tukey_weight(x: torch.Tensor, k: float) -> torch.Tensor:             # tracker.py:66-69
    w = torch.zeros_like(x)
    m = x.abs() <= k
    w[m] = (1 - (x[m] / k) ** 2) ** 2
    return w


def compute_sdf_Hg(omap: OracleMap, R_last, t_last, R_delta, t_delta, obs_xyz: np.ndarray, robust_k: float | None = 5.0,
                   no_grad: bool = False, robust_kernel: str = "huber"):
    """tracker.py:174-218.  Poses as float64 (R, t); obs_xyz (N,3) fp32 camera frame.
    Returns (H 6x6 f64, g (6,) f64, energy float) or (None, None, energy) when no_grad."""
    Rc, tc = compose(R_last, t_last, R_delta, t_delta)
    obs = torch.from_numpy(obs_xyz.astype(np.float32))
    cur = obs @ torch.from_numpy(Rc).float().t() + torch.from_numpy(tc).float().unsqueeze(0)   # motion_util.py:322-327
    if no_grad:
        sdf, std, valid = omap.get_sdf(cur.numpy())
        r = torch.from_numpy(sdf) / torch.from_numpy(std)
        JW = None
    else:
        sdf, std, valid, g = omap.get_sdf(cur.numpy(), want_grad=True)
        r = torch.from_numpy(sdf) / torch.from_numpy(std)
        G = torch.from_numpy(g)
        q = (obs @ torch.from_numpy(R_delta).float().t() + torch.from_numpy(t_delta).float().unsqueeze(0))[torch.from_numpy(valid)]
        Lt = torch.from_numpy(R_last.astype(np.float32).T)
        A = torch.mm(G, Lt)                                               # tracker.py:198
        Bm = torch.cross(q, A, dim=-1)                                    # :199
        J = torch.cat([A, Bm], dim=-1)
        JW = J
    Wf = r
    if robust_k is not None:
        w = huber_weight(r, robust_k) if robust_kernel == "huber" else tukey_weight(r, robust_k)
        Wf = Wf * w
        JW = JW * w.unsqueeze(1) if JW is not None else None
    scale = 1.0 / Wf.size(0)
    energy = (r * Wf).sum().item() * scale
    if no_grad:
        return None, None, float(energy)
    H = torch.einsum("na,nb->nab", JW, J).sum(0) * scale
    gv = (J * Wf.unsqueeze(1)).sum(0) * scale
    return H.numpy().astype(float), gv.numpy().astype(float), float(energy)


def se3_exp(xi: np.ndarray):
    """Isometry.from_twist (motion_util.py:205-229, :45-57): xi = [rho, phi] -> (R, t) with t = J_l(phi) rho."""
    rho, phi = np.asarray(xi[:3], float), np.asarray(xi[3:6], float)
    ang = np.linalg.norm(phi)

    def wedge(v):
        return np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0]], float)
    if np.isclose(ang, 0.):
        R = np.eye(3) + wedge(phi)
        Jl = np.eye(3) + 0.5 * wedge(phi)
    else:
        ax = phi / ang
        s, c = math.sin(ang), math.cos(ang)
        R = c * np.eye(3) + (1 - c) * np.outer(ax, ax) + s * wedge(ax)
        Jl = (s / ang) * np.eye(3) + (1 - s / ang) * np.outer(ax, ax) + ((1 - c) / ang) * wedge(ax)
    return R, Jl @ rho


def optimize_latent_rows(dec, latent_vecs_unique, latent_id_inv_mapping, gathered_sdf, gathered_relative_xyz, n_iters: int,
                         code_regularization: bool = False, code_reg_lambda: float = 0.0, max_sample: int = int(1.5e6)):
    """map.py:80-117 (OptimizeProcess.do_optimize) restated on torch CPU autograd: Adam(lr 1e-2) on the unique latent rows;
    per forward_model chunk (utility.py:86-118) loss = sum(-Normal(clamp(pd_sdf), pd_std).log_prob(clamp(gt))) / n_samples
    (+ lambda * sum ||row|| / n_samples when code_regularization), gradients of the chunks accumulate.  Pinned against the executed
    reference function by tests/golden/make_golden_opt.py (tests/golden/latent_opt.npz)."""
    lat = torch.as_tensor(np.asarray(latent_vecs_unique), dtype=torch.float32).clone().requires_grad_(True)
    inv = torch.as_tensor(np.asarray(latent_id_inv_mapping), dtype=torch.long)
    gt_all = torch.as_tensor(np.asarray(gathered_sdf), dtype=torch.float32)
    xyz = torch.as_tensor(np.asarray(gathered_relative_xyz), dtype=torch.float32)
    opt = torch.optim.Adam([lat], lr=1.0e-2)
    n = inv.shape[0]
    n_chunks = max(1, -(-n // max_sample))
    for _ in range(n_iters):
        opt.zero_grad()
        vec = lat[inv]
        head = 0
        for x_chunk, v_chunk in zip(torch.chunk(xyz, n_chunks), torch.chunk(vec, n_chunks)):
            sdf, std = decoder_forward(dec, v_chunk, x_chunk)
            gt = torch.clamp(gt_all[head:head + sdf.shape[0]], -0.2, 0.2)
            ll = -torch.distributions.Normal(loc=torch.clamp(sdf, -0.2, 0.2), scale=std).log_prob(gt)
            loss = ll.sum() / n
            if code_regularization:
                loss = loss + code_reg_lambda * torch.sum(torch.norm(lat, dim=1)) / n
            loss.backward(retain_graph=True)
            head += sdf.shape[0]
        opt.step()
    return lat.detach().numpy()
