"""Deterministic synthetic weights and video clips for parity tests and benchmarks.

The reference ships no checkpoint that is reachable offline (Testing/TEST_README.md:7 links to
Google Drive) and no golden vectors, so every parity check in this repo runs on *synthetic*
weights.  The weights are a pure function of (state-dict key, shape, seed): the same filler is
applied to the reference model (when the golden fixtures are generated), to the CPU oracle and to
the CUDA model, which also proves that all three agree on state-dict keys and shapes
(Testing/model/pspnet/td4_psp18.py:232-240 loads with strict=True).

BatchNorm running statistics and the LayerNorm affine are randomised on purpose so that BN folding
and LN bugs are visible (constructor defaults would make them identities).
"""
from __future__ import annotations

import hashlib
import math
import re

import torch


_RESIDUAL_TAIL_BN = re.compile(r"pretrained\d*\.layer\d+\.\d+\.(bn2|bn3)\.weight$")
# TD2-FANet (td2_fa.py:334-349): the linear-attention branch sums over all pixels of the map, so its BatchNorm sees
# inputs whose scale grows with the image; a trained network absorbs that in the running variance, the synthetic
# one in a small gamma.  `smooth` feeds the 256-channel map the Encoding projections read.
_FA_LATERAL_BN = re.compile(r"ffm_\d+_\d\.latlayer3\.bn\.weight$")
_FA_SMOOTH_BN = re.compile(r"ffm_\d+_\d\.smooth\.bn\.weight$")


def _gen_for(key: str, seed: int) -> torch.Generator:
    digest = hashlib.sha256(f"{seed}:{key}".encode()).digest()
    g = torch.Generator(device="cpu")
    g.manual_seed(int.from_bytes(digest[:8], "little") & 0x7FFF_FFFF_FFFF_FFFF)
    return g


def synth_tensor(key: str, ref: torch.Tensor, seed: int = 0, tame: bool = True) -> torch.Tensor:
    """Value for state-dict entry `key` (same shape/dtype as `ref`).  tame=False drops the two adjustments that keep a
    random network as calm as a trained one (small residual-tail gammas, 0.25 gain on the second Q/K projection):
    activations then grow block by block and the softmax turns one-hot -- the stress case of the range guard."""
    g = _gen_for(key, seed)
    shape = tuple(ref.shape)
    leaf = key.rsplit(".", 1)[-1]
    if leaf == "num_batches_tracked":
        return torch.zeros(shape, dtype=ref.dtype)
    if leaf == "running_mean":
        return torch.randn(shape, generator=g) * 0.1
    if leaf == "running_var":
        return torch.rand(shape, generator=g) + 0.5
    if ref.dim() == 4:  # conv weight [cout, cin, kh, kw]
        fan_in = shape[1] * shape[2] * shape[3]
        if (".w_qs.1." in key or ".w_ks.1." in key) and tame:
            gain = 0.25  # keeps q.k/8 at a few units so the softmax is neither flat nor one-hot
        elif ".w_qs." in key or ".w_ks." in key or ".w_vs." in key or ".fc." in key:
            gain = 1.0
        else:
            gain = 2.0
        return torch.randn(shape, generator=g) * math.sqrt(gain / fan_in)
    if ref.dim() == 2 and leaf == "weight" and ".ln." not in key:  # resnet fc (unused by forward)
        return torch.randn(shape, generator=g) * math.sqrt(1.0 / shape[1])
    if leaf == "weight":  # BN gamma [C] or LN gamma [H8, W8]
        gamma = torch.rand(shape, generator=g) + 0.5
        if not tame:
            pass
        elif _RESIDUAL_TAIL_BN.search(key):
            # Last BN of a residual block: keep the branch small so that the trunk does not double
            # its variance at every block (a trained network is calm; a random one is not).
            gamma = gamma * 0.3
        elif _FA_LATERAL_BN.search(key):
            gamma = gamma * 0.02
        elif _FA_SMOOTH_BN.search(key):
            gamma = gamma * 0.25
        return gamma
    if leaf == "bias":
        return torch.randn(shape, generator=g) * 0.1
    raise KeyError(f"synth_tensor: unclassified state-dict key {key!r} shape {shape}")


def synth_state_dict(template: dict, seed: int = 0, tame: bool = True) -> dict:
    """Fill every entry of a state-dict template (key -> tensor) deterministically."""
    return {k: synth_tensor(k, v, seed, tame).to(v.dtype) for k, v in template.items()}


_MEAN = torch.tensor([0.485, 0.456, 0.406]).view(1, 3, 1, 1)
_STD = torch.tensor([0.229, 0.224, 0.225]).view(1, 3, 1, 1)


def synth_clip(n_frames: int, height: int, width: int, batch: int = 1, clip_id: int = 0):
    """Seeded Cityscapes-shaped clip: list of fp32 NCHW frames normalised like
    Testing/dataloader.py:52-53,66-71 ((x/255 - mean)/std).  Frame t is the base image rolled by
    (t, 2t) pixels plus N(0, 2) sensor noise, so consecutive frames are correlated as in video."""
    g = torch.Generator(device="cpu")
    g.manual_seed(1234 + clip_id)
    lh, lw = max(height // 8, 2), max(width // 8, 2)
    low = torch.randint(0, 256, (batch, 3, lh, lw), generator=g).float()
    base = torch.nn.functional.interpolate(low, size=(height, width), mode="bicubic",
                                           align_corners=False).clamp_(0, 255)
    frames = []
    for t in range(n_frames):
        img = torch.roll(base, shifts=(t, 2 * t), dims=(2, 3))
        img = (img + torch.randn(img.shape, generator=g) * 2.0).clamp_(0, 255).round_()
        frames.append(((img / 255.0 - _MEAN) / _STD).contiguous())
    return frames
