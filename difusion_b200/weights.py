"""Fold the reference checkpoint into the fp32 blobs of include/difusion_b200.h and upload/prepare them.

  decoder: W_k = g_k * v_k / ||v_k||_row   (torch weight_norm, reference network/di_decoder.py:36-40)
  encoder: eval-mode BatchNorm folded into the 1x1 convs (reference utils/pt_util.py:37-116)

Accepts the reference's state_dict key names (``lin{k}.weight_g/_v`` or the newer parametrization names,
``mlp.layer{k}.conv.weight`` ...).  Host-side numpy only; the device-side expansion is ``dif_prepare_*``.
"""
from __future__ import annotations

import numpy as np

DECODER_BLOB_FLOATS = 128 * 32 + 128 + 128 * 128 + 128 + 96 * 128 + 96 + 128 * 128 + 128 + 128 + 1 + 128 + 1
ENCODER_BLOB_FLOATS = 32 * 6 + 32 + 64 * 32 + 64 + 256 * 64 + 256 + 29 * 256 + 29


def _np(v):
    return v.detach().cpu().numpy() if hasattr(v, "detach") else np.asarray(v)


def _wn(sd, k):
    for g_name, v_name in ((f"lin{k}.weight_g", f"lin{k}.weight_v"),
                           (f"lin{k}.parametrizations.weight.original0", f"lin{k}.parametrizations.weight.original1")):
        if g_name in sd:
            g, v = _np(sd[g_name]).astype(np.float32), _np(sd[v_name]).astype(np.float32)
            norm = np.sqrt((v.astype(np.float32) ** 2).sum(axis=1, keepdims=True, dtype=np.float32))
            return (v * (g / norm)).astype(np.float32)
    return _np(sd[f"lin{k}.weight"]).astype(np.float32)          # a decoder saved without weight-norm


def fold_decoder(sd: dict) -> np.ndarray:
    parts = []
    shapes = [(128, 32), (128, 128), (96, 128), (128, 128)]
    for k, shp in enumerate(shapes):
        W = _wn(sd, k)
        if W.shape != shp:
            raise ValueError(f"decoder lin{k}: expected {shp}, got {W.shape} (only the shipped 29+3 -> 4x128 architecture is built)")
        parts += [W.ravel(), _np(sd[f"lin{k}.bias"]).astype(np.float32).ravel()]
    parts += [_wn(sd, 4).ravel(), _np(sd["lin4.bias"]).astype(np.float32).ravel(),
              _np(sd["uncertainty_layer.weight"]).astype(np.float32).ravel(), _np(sd["uncertainty_layer.bias"]).astype(np.float32).ravel()]
    blob = np.concatenate(parts).astype(np.float32)
    assert blob.size == DECODER_BLOB_FLOATS
    return blob


def fold_encoder(sd: dict, eps: float = 1e-5) -> np.ndarray:
    parts = []
    for k, shp in enumerate([(32, 6), (64, 32), (256, 64)]):
        W = _np(sd[f"mlp.layer{k}.conv.weight"]).astype(np.float32).reshape(shp)
        bn = f"mlp.layer{k}.normlayer.bn."
        scale = _np(sd[bn + "weight"]).astype(np.float32) / np.sqrt(_np(sd[bn + "running_var"]).astype(np.float32) + np.float32(eps))
        parts += [(W * scale[:, None]).ravel(), (_np(sd[bn + "bias"]).astype(np.float32) - _np(sd[bn + "running_mean"]).astype(np.float32) * scale).ravel()]
    parts += [_np(sd["mlp.layer3.conv.weight"]).astype(np.float32).reshape(29, 256).ravel(), _np(sd["mlp.layer3.conv.bias"]).astype(np.float32).ravel()]
    blob = np.concatenate(parts).astype(np.float32)
    assert blob.size == ENCODER_BLOB_FLOATS
    return blob


def load_npz_state(path):
    """(decoder_state, encoder_state) from the raw-checkpoint npz written by tests/golden/make_golden.py."""
    z = np.load(path)
    return ({k[4:]: z[k] for k in z.files if k.startswith("dec.")}, {k[4:]: z[k] for k in z.files if k.startswith("enc.")})


class PreparedNetworks:
    """Device-resident prepared weights for the kernels (one per device)."""

    def __init__(self, decoder_state: dict | None, encoder_state: dict | None, device):
        import torch
        from . import _lib
        L = _lib.lib()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.DifusionLibraryError("difusion_b200 runs on CUDA devices only (no CPU fallback)")
        st = _lib.stream_ptr(self.device)
        with torch.cuda.device(self.device):
            self.decoder = None
            if decoder_state is not None:
                blob = torch.from_numpy(fold_decoder(decoder_state)).to(self.device)
                self.decoder = torch.empty(L.dif_decoder_prepared_bytes(), dtype=torch.uint8, device=self.device)
                _lib.check(L.dif_prepare_decoder(blob.data_ptr(), self.decoder.data_ptr(), st), "dif_prepare_decoder")
                self.decoder_blob = blob
            self.encoder = None
            if encoder_state is not None:
                eblob = torch.from_numpy(fold_encoder(encoder_state)).to(self.device)
                self.encoder = torch.empty(L.dif_encoder_prepared_bytes(), dtype=torch.uint8, device=self.device)
                _lib.check(L.dif_prepare_encoder(eblob.data_ptr(), self.encoder.data_ptr(), st), "dif_prepare_encoder")
                self.encoder_blob = eblob
