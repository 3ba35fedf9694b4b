"""Host-side mirror of the reference's ``network/utility.py`` for the hot path (same names, arguments and errors):
``Networks``, ``load_model``, ``forward_model``, ``get_samples``, ``groupby_reduce``.  All arithmetic runs in
libdifusion_b200.so; there is no torch fallback.
"""
from __future__ import annotations

import json
import math
from pathlib import Path

import torch

from .. import _lib, weights

LATENT_DIM = 29


class _Net:
    """A network described only by its (reference-format) state dict; evaluated by the CUDA library."""

    def __init__(self, state: dict):
        self._state = dict(state)
        self.training = False

    def state_dict(self):
        return self._state

    def eval(self):
        self.training = False
        return self

    def cuda(self, *a, **k):
        return self

    def to(self, *a, **k):
        return self


class Decoder(_Net):
    """network/di_decoder.py Model: callable on an (n, 32) input like the reference module."""

    def __call__(self, network_input: torch.Tensor):
        return _decode_autograd(self, network_input[:, :LATENT_DIM], network_input[:, LATENT_DIM:])


class Encoder(_Net):
    """network/di_encoder.py Model in 'cnp' mode: (S, 6) -> (S, 29)."""

    def __call__(self, xyzn: torch.Tensor):
        prep = prepared_for(Networks.of(encoder=self), xyzn.device)
        x = xyzn.detach().contiguous().float()
        out = torch.empty((x.size(0), LATENT_DIM), dtype=torch.float32, device=x.device)
        _lib.check(_lib.lib().dif_encode(prep.encoder.data_ptr(), _lib.ptr(x), x.size(0), _lib.ptr(out), _lib.stream_ptr(x.device)), "dif_encode")
        return out


class Networks:                                          # reference utility.py:10-19
    def __init__(self):
        self.decoder = None
        self.encoder = None
        self._prepared = {}

    @staticmethod
    def of(decoder=None, encoder=None):
        n = Networks()
        n.decoder, n.encoder = decoder, encoder
        return n

    def eval(self):
        if self.encoder is not None:
            self.encoder.eval()
        if self.decoder is not None:
            self.decoder.eval()


_PREP_CACHE = {}


def prepared_for(model, device) -> weights.PreparedNetworks:
    """Prepared (device-resident) weights for a Networks-like object holding reference-format modules or _Net objects."""
    device = torch.device(device)
    if device.type == "cuda" and device.index is None:
        device = torch.device("cuda", torch.cuda.current_device())
    key = (id(model.decoder), id(model.encoder), str(device))
    p = _PREP_CACHE.get(key)
    if p is None:
        dec = model.decoder.state_dict() if model.decoder is not None else None
        enc = model.encoder.state_dict() if model.encoder is not None else None
        p = weights.PreparedNetworks(dec, enc, device)
        _PREP_CACHE[key] = p
        p._keepalive = (model.decoder, model.encoder)
    return p


def load_model(training_hyper_path: str, use_epoch: int = -1, device="cuda"):
    """reference utility.py:22-58: hyper.json next to model_{epoch}.pth.tar / encoder_{epoch}.pth.tar.
    Also accepts a raw-checkpoint ``.npz`` (tests/golden/weights.npz).  Returns (Networks, args)."""
    from types import SimpleNamespace
    path = Path(training_hyper_path)
    model = Networks()
    if path.suffix == ".npz":
        dec, enc = weights.load_npz_state(path)
        model.decoder, model.encoder = Decoder(dec), Encoder(enc)
        return model, SimpleNamespace(code_length=LATENT_DIM)
    if path.suffix != ".json":
        raise NotImplementedError("only trained checkpoints (hyper.json) are supported")
    args = SimpleNamespace(**json.load(open(path)))
    ckpts = {int(str(t).split("model_")[-1].split(".pth")[0]): t for t in path.parent.glob("model_*.pth.tar")}
    assert use_epoch in ckpts.keys(), f"{use_epoch} not found in {sorted(list(ckpts.keys()))}"
    args.checkpoint = ckpts[use_epoch]
    model.decoder = Decoder(torch.load(args.checkpoint, map_location="cpu")["model_state"])
    if getattr(args, "encoder_name", None) is not None:
        model.encoder = Encoder(torch.load(path.parent / f"encoder_{use_epoch}.pth.tar", map_location="cpu")["model_state"])
    return model, args


class _DecodeFn(torch.autograd.Function):
    """sdf, std = decoder(latent, xyz) with analytic d/dxyz from the fused forward+backward kernel."""

    @staticmethod
    def forward(ctx, xyz, latent, prep):
        L = _lib.lib()
        n = xyz.size(0)
        dev = xyz.device
        x = xyz.detach().contiguous().float()
        lat = latent.detach().contiguous().float()
        sdf = torch.empty(n, dtype=torch.float32, device=dev)
        std = torch.empty(n, dtype=torch.float32, device=dev)
        need = xyz.requires_grad
        g = torch.empty((n, 3), dtype=torch.float32, device=dev) if need else None
        _lib.check(L.dif_decode(prep.decoder.data_ptr(), _lib.ptr(lat), LATENT_DIM, None, _lib.ptr(x), n, None, 1.0, _lib.ptr(sdf), _lib.ptr(std),
                                _lib.ptr(g), None, _lib.stream_ptr(dev)), "dif_decode")
        ctx.prep = prep
        ctx.save_for_backward(x, lat, g if need else torch.empty(0, device=dev))
        ctx.mark_non_differentiable()
        return sdf, std

    @staticmethod
    def backward(ctx, g_sdf, g_std):
        x, lat, dsdf = ctx.saved_tensors
        grad = None
        if dsdf.numel() > 0:
            grad = g_sdf.unsqueeze(-1) * dsdf
            if g_std is not None and bool((g_std != 0).any()):
                L = _lib.lib()
                n = x.size(0)
                s0 = torch.empty(n, dtype=torch.float32, device=x.device)
                s1 = torch.empty(n, dtype=torch.float32, device=x.device)
                g0 = torch.empty((n, 3), dtype=torch.float32, device=x.device)
                g1 = torch.empty((n, 3), dtype=torch.float32, device=x.device)
                _lib.check(L.dif_decode(ctx.prep.decoder.data_ptr(), _lib.ptr(lat), LATENT_DIM, None, _lib.ptr(x), n, None, 1.0, _lib.ptr(s0),
                                        _lib.ptr(s1), _lib.ptr(g0), _lib.ptr(g1), _lib.stream_ptr(x.device)), "dif_decode")
                grad = grad + g_std.unsqueeze(-1) * g1
        return grad, None, None


def _decode_autograd(decoder, latent, xyz):
    prep = prepared_for(Networks.of(decoder=decoder), xyz.device)
    sdf, std = _DecodeFn.apply(xyz, latent, prep)
    return sdf.unsqueeze(-1), std.unsqueeze(-1)


def forward_model(model, network_input: torch.Tensor = None, latent_input: torch.Tensor = None, xyz_input: torch.Tensor = None,
                  loss_func=None, max_sample: int = 2 ** 32, no_detach: bool = False, verbose: bool = False):
    """reference utility.py:61-126.  ``model`` is a decoder (``Decoder`` or anything with the reference state_dict).
    Returns [sdf (N,1), std (N,1)].  Gradients flow to ``xyz_input`` when no_detach=True (the tracker's use);
    ``loss_func`` (latent optimisation, disabled in the reference's shipped configuration) is not supported."""
    if loss_func is not None:
        raise NotImplementedError("forward_model(loss_func=...) belongs to the disabled latent-optimisation path (map.py:456-516)")
    if latent_input is not None and xyz_input is not None:
        assert network_input is None
    else:
        assert network_input is not None and network_input.ndimension() == 2
        latent_input, xyz_input = network_input[:, :LATENT_DIM], network_input[:, LATENT_DIM:]
    n = latent_input.size(0)
    assert n > 0                                                    # the reference asserts via torch.chunk (utility.py:82-88)
    n_chunks = math.ceil(n / max_sample)
    assert not no_detach or n_chunks == 1
    sdf, std = _decode_autograd(model, latent_input, xyz_input)
    if not no_detach:
        sdf, std = sdf.detach(), std.detach()
    return [sdf, std]


def get_samples(r: int, device: torch.device, a: float = 0.0, b: float = None):
    """reference utility.py:129-149 (the kernels generate this lattice on the fly; kept for API parity)."""
    idx = torch.arange(0, r ** 3, 1, device=device, dtype=torch.long)
    r = int(r)
    if b is None:
        b = 1. - 1. / r
    vsize = (b - a) / (r - 1)
    s = torch.zeros(r ** 3, 3, device=device, dtype=torch.float32)
    s[:, 0] = (idx // (r * r)) * vsize + a
    s[:, 1] = ((idx // r) % r) * vsize + a
    s[:, 2] = (idx % r) * vsize + a
    return s


def groupby_reduce(sample_indexer: torch.Tensor, sample_values: torch.Tensor, op: str = "max"):
    """reference utility.py:186-206."""
    from ..system.ext import groupby_sum
    C = int(sample_indexer.max()) + 1
    assert sample_indexer.size(0) == sample_values.size(0), "Indexer and Values must agree on sample count!"
    if op == "mean":
        s, c = groupby_sum(sample_values, sample_indexer, C)
        return s / c.unsqueeze(-1)
    elif op == "sum":
        return groupby_sum(sample_values, sample_indexer, C)[0]
    raise NotImplementedError
