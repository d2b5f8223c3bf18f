"""SNAC-24k decode-path weights: checkpoint I/O, seeded random init, weight-norm folding.

The reference loads ``SNAC.from_pretrained(path_or_repo)`` at import time
(``/root/reference/Morpheus_Client/tts_engine/speechpipe.py:38-43``) and lets the
third-party model recompute ``w = g*v/||v||`` in every forward.  Here the
state dict is folded ONCE on the host into plain fp32 arrays in PyTorch's native
layouts; the C-ABI library re-packs them for its kernels (``snacb_load_weights``).

State-dict keys follow the published ``snac`` package so a real
``pytorch_model.bin`` loads unchanged; both weight-norm spellings are accepted
(``weight_g``/``weight_v`` and ``parametrizations.weight.original0``/``original1``).
"""
from __future__ import annotations

import json
import math
import os
from typing import Dict, List, Tuple

import torch

SNAC24K_CONFIG = {
    "sampling_rate": 24000,
    "encoder_dim": 48,
    "encoder_rates": [2, 4, 8, 8],
    "latent_dim": None,
    "decoder_dim": 1024,
    "decoder_rates": [8, 8, 4, 2],
    "attn_window_size": None,
    "codebook_size": 4096,
    "codebook_dim": 8,
    "vq_strides": [4, 2, 1],
    "noise": True,
    "depthwise": True,
}
LATENT = 768
DECODER_DIM = 1024
RATES = (8, 8, 4, 2)
VQ_STRIDES = (4, 2, 1)
CODEBOOK_SIZE = 4096
CODEBOOK_DIM = 8
DILATIONS = (1, 3, 9)

_ALIASES = (
    (".parametrizations.weight.original0", ".weight_g"),
    (".parametrizations.weight.original1", ".weight_v"),
)


def normalise_keys(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    out = {}
    for k, v in sd.items():
        for new, old in _ALIASES:
            if k.endswith(new):
                k = k[: -len(new)] + old
        out[k] = v
    return out


def check_config(cfg: dict) -> None:
    """Only the snac_24khz geometry has kernels; anything else fails loudly."""
    for key in ("decoder_dim", "decoder_rates", "codebook_size", "codebook_dim", "vq_strides", "noise", "depthwise"):
        if key in cfg and cfg[key] != SNAC24K_CONFIG[key]:
            raise ValueError(f"unsupported SNAC config: {key}={cfg[key]!r} (kernels are built for snac_24khz)")
    if cfg.get("attn_window_size") is not None:
        raise ValueError("unsupported SNAC config: attention layer present (snac_24khz has none)")
    latent = cfg.get("latent_dim") or cfg.get("encoder_dim", 48) * 2 ** len(cfg.get("encoder_rates", [2, 4, 8, 8]))
    if latent != LATENT:
        raise ValueError(f"unsupported SNAC config: latent_dim={latent}")


# --------------------------------------------------------------------------- layer table
def decoder_layout() -> List[Tuple[str, str, tuple]]:
    """(state-dict prefix, kind, shape info) for every parametrised layer on the decode path."""
    rows: List[Tuple[str, str, tuple]] = []
    for i in range(3):
        rows.append((f"quantizer.quantizers.{i}.codebook", "embedding", (CODEBOOK_SIZE, CODEBOOK_DIM)))
        rows.append((f"quantizer.quantizers.{i}.out_proj", "conv", (LATENT, CODEBOOK_DIM, 1, True)))
    rows.append(("decoder.model.0", "conv", (LATENT, 1, 7, True)))
    rows.append(("decoder.model.1", "conv", (DECODER_DIM, LATENT, 1, True)))
    c = DECODER_DIM
    for b, s in enumerate(RATES):
        p = f"decoder.model.{2 + b}.block"
        rows.append((f"{p}.0", "snake", (c,)))
        rows.append((f"{p}.1", "convT", (c, c // 2, 2 * s)))
        rows.append((f"{p}.2.linear", "conv", (c // 2, c // 2, 1, False)))
        for r in range(3):
            q = f"{p}.{3 + r}.block"
            rows.append((f"{q}.0", "snake", (c // 2,)))
            rows.append((f"{q}.1", "conv", (c // 2, 1, 7, True)))
            rows.append((f"{q}.2", "snake", (c // 2,)))
            rows.append((f"{q}.3", "conv", (c // 2, c // 2, 1, True)))
        c //= 2
    rows.append(("decoder.model.6", "snake", (c,)))
    rows.append(("decoder.model.7", "conv", (1, c, 7, True)))
    return rows


ENCODER_DIM = 48
ENCODER_RATES = (2, 4, 8, 8)


def encoder_layout() -> List[Tuple[str, str, tuple]]:
    """Parametrised layers of the encode direction: encoder + the quantizers' in-projections."""
    rows: List[Tuple[str, str, tuple]] = [("encoder.block.0", "conv", (ENCODER_DIM, 1, 7, True))]
    c = ENCODER_DIM
    for b, s in enumerate(ENCODER_RATES):
        p = f"encoder.block.{1 + b}.block"
        for r in range(3):
            q = f"{p}.{r}.block"
            rows.append((f"{q}.0", "snake", (c,)))
            rows.append((f"{q}.1", "conv", (c, 1, 7, True)))
            rows.append((f"{q}.2", "snake", (c,)))
            rows.append((f"{q}.3", "conv", (c, c, 1, True)))
        rows.append((f"{p}.3", "snake", (c,)))
        rows.append((f"{p}.4", "conv", (2 * c, c, 2 * s, True)))
        c *= 2
    rows.append(("encoder.block.5", "conv", (c, 1, 7, True)))
    for i in range(3):
        rows.append((f"quantizer.quantizers.{i}.in_proj", "conv", (CODEBOOK_DIM, LATENT, 1, True)))
    return rows


def random_encoder_state_dict(seed: int = 0, variant: str = "default") -> Dict[str, torch.Tensor]:
    """Seeded random-init weights of the ENCODE direction, drawn from generators of their own so the decode-path
    weights of ``random_state_dict`` (and every golden vector made from them) are unchanged.  Merge the two dicts for a
    full model."""
    g = torch.Generator(device="cpu").manual_seed(1000003 + seed)
    g1 = torch.Generator(device="cpu").manual_seed(1000003 + seed + 1)
    w1 = variant == "w1"
    if variant not in ("default", "w1"):
        raise ValueError(variant)
    sd: Dict[str, torch.Tensor] = {}
    for prefix, kind, info in encoder_layout():
        if kind == "snake":
            a = torch.ones(1, info[0], 1)
            if w1:
                a = 0.5 + torch.rand(a.shape, generator=g1)
            sd[f"{prefix}.alpha"] = a
        else:
            cout, cin_g, k, _ = info
            bound = 1.0 / math.sqrt(cin_g * k)
            v = (torch.rand((cout, cin_g, k), generator=g) * 2 - 1) * bound
            gain = v.flatten(1).norm(dim=1).reshape(-1, 1, 1)
            if w1:
                gain = gain * (0.8 + 0.4 * torch.rand(gain.shape, generator=g1))
            sd[f"{prefix}.weight_g"], sd[f"{prefix}.weight_v"] = gain, v
            sd[f"{prefix}.bias"] = (torch.rand((cout,), generator=g) * 2 - 1) * bound
    return sd


# --------------------------------------------------------------------------- random init
def random_state_dict(seed: int = 0, variant: str = "default") -> Dict[str, torch.Tensor]:
    """Seeded random-init weights of the snac_24khz decode path (no network for checkpoints).

    Statistics follow PyTorch's default conv init (``U(-1/sqrt(fan_in), 1/sqrt(fan_in))`` for
    weight and bias, ``g = ||v||``, Snake ``alpha = 1``, codebook ``N(0,1)``).  ``variant="w1"``
    additionally perturbs ``g <- g*U(0.8,1.2)`` and ``alpha <- U(0.5,1.5)`` so the weight-norm
    gains and per-channel alphas are exercised (SURVEY 8(d) "Variant W1").
    """
    g = torch.Generator(device="cpu").manual_seed(seed)
    g1 = torch.Generator(device="cpu").manual_seed(seed + 1)
    w1 = variant == "w1"
    if variant not in ("default", "w1"):
        raise ValueError(variant)

    def uni(shape, bound):
        return (torch.rand(shape, generator=g) * 2 - 1) * bound

    sd: Dict[str, torch.Tensor] = {}
    for prefix, kind, info in decoder_layout():
        if kind == "embedding":
            sd[f"{prefix}.weight"] = torch.randn(info, generator=g)
        elif kind == "snake":
            a = torch.ones(1, info[0], 1)
            if w1:
                a = 0.5 + torch.rand(a.shape, generator=g1)
            sd[f"{prefix}.alpha"] = a
        elif kind == "conv":
            cout, cin_g, k, bias = info
            bound = 1.0 / math.sqrt(cin_g * k)
            v = uni((cout, cin_g, k), bound)
            gain = v.flatten(1).norm(dim=1).reshape(-1, 1, 1)
            if w1:
                gain = gain * (0.8 + 0.4 * torch.rand(gain.shape, generator=g1))
            sd[f"{prefix}.weight_g"], sd[f"{prefix}.weight_v"] = gain, v
            if bias:
                sd[f"{prefix}.bias"] = uni((cout,), bound)
        elif kind == "convT":
            cin, cout, k = info
            bound = 1.0 / math.sqrt(cout * k)  # torch computes fan_in from dim 1 of [cin, cout, k]
            v = uni((cin, cout, k), bound)
            gain = v.flatten(1).norm(dim=1).reshape(-1, 1, 1)
            if w1:
                gain = gain * (0.8 + 0.4 * torch.rand(gain.shape, generator=g1))
            sd[f"{prefix}.weight_g"], sd[f"{prefix}.weight_v"] = gain, v
            sd[f"{prefix}.bias"] = uni((cout,), bound)
    return sd


def save_checkpoint(path: str, sd: Dict[str, torch.Tensor]) -> None:
    """Write the ``config.json`` + ``pytorch_model.bin`` layout ``from_pretrained`` expects."""
    os.makedirs(path, exist_ok=True)
    with open(os.path.join(path, "config.json"), "w") as f:
        json.dump(SNAC24K_CONFIG, f)
    torch.save(dict(sd), os.path.join(path, "pytorch_model.bin"))


def load_checkpoint(path: str) -> Dict[str, torch.Tensor]:
    with open(os.path.join(path, "config.json")) as f:
        check_config(json.load(f))
    return torch.load(os.path.join(path, "pytorch_model.bin"), map_location="cpu", weights_only=True)


# --------------------------------------------------------------------------- folding
def _fold(sd, prefix: str) -> torch.Tensor:
    """``w = g * v / ||v||`` with the norm over every dim but 0 (weight_norm ``dim=0``).

    For ConvTranspose1d the stored weight is ``[Cin, Cout, k]`` so this is per INPUT channel.
    """
    if f"{prefix}.weight" in sd:  # already-plain weight
        return sd[f"{prefix}.weight"].float()
    v = sd[f"{prefix}.weight_v"].float()
    gain = sd[f"{prefix}.weight_g"].float()
    return torch._weight_norm(v, gain, 0)


class FoldedWeights:
    """Folded fp32 tensors (contiguous, CPU) in the order/layouts ``snacb_weights`` declares."""

    def __init__(self, sd: Dict[str, torch.Tensor]):
        sd = normalise_keys(sd)
        t: Dict[str, torch.Tensor] = {}

        def put(name, x):
            t[name] = x.detach().to(torch.float32).contiguous().cpu()

        for i in range(3):
            q = f"quantizer.quantizers.{i}"
            put(f"codebook{i}", sd[f"{q}.codebook.weight"])                    # [4096, 8]
            put(f"outproj_w{i}", _fold(sd, f"{q}.out_proj").reshape(LATENT, CODEBOOK_DIM))
            put(f"outproj_b{i}", sd[f"{q}.out_proj.bias"])
        put("head_dw_w", _fold(sd, "decoder.model.0").reshape(LATENT, 7))
        put("head_dw_b", sd["decoder.model.0.bias"])
        put("head_pw_w", _fold(sd, "decoder.model.1").reshape(DECODER_DIM, LATENT))
        put("head_pw_b", sd["decoder.model.1.bias"])
        c = DECODER_DIM
        for b, s in enumerate(RATES):
            p = f"decoder.model.{2 + b}.block"
            put(f"b{b}_alpha", sd[f"{p}.0.alpha"].reshape(c))
            put(f"b{b}_convt_w", _fold(sd, f"{p}.1"))                          # [Cin, Cout, 2s]
            put(f"b{b}_convt_b", sd[f"{p}.1.bias"])
            put(f"b{b}_noise_w", _fold(sd, f"{p}.2.linear").reshape(c // 2, c // 2))
            for r in range(3):
                q = f"{p}.{3 + r}.block"
                put(f"b{b}_r{r}_alpha1", sd[f"{q}.0.alpha"].reshape(c // 2))
                put(f"b{b}_r{r}_dw_w", _fold(sd, f"{q}.1").reshape(c // 2, 7))
                put(f"b{b}_r{r}_dw_b", sd[f"{q}.1.bias"])
                put(f"b{b}_r{r}_alpha2", sd[f"{q}.2.alpha"].reshape(c // 2))
                put(f"b{b}_r{r}_pw_w", _fold(sd, f"{q}.3").reshape(c // 2, c // 2))
                put(f"b{b}_r{r}_pw_b", sd[f"{q}.3.bias"])
            c //= 2
        put("tail_alpha", sd["decoder.model.6.alpha"].reshape(c))
        put("tail_w", _fold(sd, "decoder.model.7").reshape(c, 7))
        put("tail_b", sd["decoder.model.7.bias"])
        # encode direction (N4), when the checkpoint carries it
        self.has_encoder = "encoder.block.0.bias" in sd
        if self.has_encoder:
            put("enc_in_w", _fold(sd, "encoder.block.0").reshape(ENCODER_DIM, 7))
            put("enc_in_b", sd["encoder.block.0.bias"])
            c = ENCODER_DIM
            for b, s in enumerate(ENCODER_RATES):
                p = f"encoder.block.{1 + b}.block"
                for r in range(3):
                    q = f"{p}.{r}.block"
                    put(f"enc{b}_r{r}_alpha1", sd[f"{q}.0.alpha"].reshape(c))
                    put(f"enc{b}_r{r}_dw_w", _fold(sd, f"{q}.1").reshape(c, 7))
                    put(f"enc{b}_r{r}_dw_b", sd[f"{q}.1.bias"])
                    put(f"enc{b}_r{r}_alpha2", sd[f"{q}.2.alpha"].reshape(c))
                    put(f"enc{b}_r{r}_pw_w", _fold(sd, f"{q}.3").reshape(c, c))
                    put(f"enc{b}_r{r}_pw_b", sd[f"{q}.3.bias"])
                put(f"enc{b}_alpha", sd[f"{p}.3.alpha"].reshape(c))
                put(f"enc{b}_down_w", _fold(sd, f"{p}.4"))                      # [2C, C, 2s]
                put(f"enc{b}_down_b", sd[f"{p}.4.bias"])
                c *= 2
            put("enc_out_dw_w", _fold(sd, "encoder.block.5").reshape(LATENT, 7))
            put("enc_out_dw_b", sd["encoder.block.5.bias"])
            for i in range(3):
                q = f"quantizer.quantizers.{i}"
                put(f"inproj_w{i}", _fold(sd, f"{q}.in_proj").reshape(CODEBOOK_DIM, LATENT))
                put(f"inproj_b{i}", sd[f"{q}.in_proj.bias"])
        self.tensors = t

    def num_params(self) -> int:
        return sum(x.numel() for x in self.tensors.values())
