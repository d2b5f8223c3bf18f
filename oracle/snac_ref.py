"""Oracle: restatement of the SNAC-24k *decode* algorithm in plain PyTorch (CPU).

TEST INFRASTRUCTURE ONLY - see ``oracle/__init__.py``.  PARITY UNPINNED: the
third-party package this restates (PyPI ``snac`` 1.2.x, pinned by the reference at
``/root/reference/requirements.txt:8`` and called only at
``Morpheus_Client/tts_engine/speechpipe.py:1,43,49,118``) is absent from the
reference tree and from this image, and the reference holds no golden vector.
The restatement is cross-checked structurally (``tests/test_oracle_model.py``):
decoder parameter count 13 012 418, decode-path parameter count 13 133 762,
output length 2048 samples per 7-token frame.

What the published algorithm computes (config of ``hubertsiuzdak/snac_24khz``:
latent 768, decoder_dim 1024, decoder_rates [8,8,4,2], codebook 4096x8,
vq_strides [4,2,1], noise=True, depthwise=True, no attention):

``decode(codes)``:
  z   = sum_i repeat_interleave(out_proj_i(codebook_i[codes_i]^T), stride_i)      (quantizer.from_codes)
  y   = decoder(z)                                                               (nn.Sequential)
decoder = [ WN dwConv k7 (768) , WN Conv1x1 768->1024 ,
            4 x DecoderBlock( Snake, WN ConvTranspose1d k=2s stride=s pad=ceil(s/2),
                              NoiseBlock, ResidualUnit d=1, d=3, d=9 ),
            Snake(64), WN Conv k7 64->1, Tanh ]
Snake(x)        = x + (alpha + 1e-9)^-1 * sin(alpha x)^2        (alpha per channel)
NoiseBlock(x)   = x + randn([B,1,T]) * WNConv1x1_nobias(x)
ResidualUnit(x) = x + WNConv1x1( Snake( WN dwConv k7 dil d ( Snake(x) ) ) )
weight_norm     : w = g * v / ||v||  (norm over all dims but 0; recomputed every
                  forward, which is what the reference does - SURVEY K10).

The one deliberate extension over the reference behaviour is *noise injection*:
``NoiseBlock`` draws from the global RNG inside ``decode`` (two seeds differ by
>20 000 LSB under random init), so parity is only definable when oracle and
kernel see the same noise.  ``SNAC.noise`` selects: ``"randn"`` (reference
behaviour), ``"off"`` (zeros) or a list of four tensors ``[B,1,T_b]``.
"""
from __future__ import annotations

import json
import math
import os
from typing import Dict, List, Optional, Sequence, Union

import torch
import torch.nn as nn
import torch.nn.functional as F

CONFIG_24KHZ = {
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


# --------------------------------------------------------------------------- layers
class _WNBase(nn.Module):
    """A conv whose weight is stored as (g, v) and normalised on every forward."""

    def _w(self) -> torch.Tensor:
        return torch._weight_norm(self.weight_v, self.weight_g, 0)


class WNConv1d(_WNBase):
    def __init__(self, cin, cout, kernel_size, padding=0, dilation=1, groups=1, bias=True, stride=1):
        super().__init__()
        conv = nn.Conv1d(cin, cout, kernel_size, padding=padding, dilation=dilation, groups=groups, bias=bias)
        v = conv.weight.detach()
        self.weight_g = nn.Parameter(v.flatten(1).norm(dim=1).reshape(-1, 1, 1).clone())
        self.weight_v = nn.Parameter(v.clone())
        self.bias = nn.Parameter(conv.bias.detach().clone()) if bias else None
        self.padding, self.dilation, self.groups, self.stride = padding, dilation, groups, stride

    def forward(self, x):
        return F.conv1d(x, self._w(), self.bias, self.stride, self.padding, self.dilation, self.groups)


class WNConvTranspose1d(_WNBase):
    def __init__(self, cin, cout, kernel_size, stride, padding, output_padding):
        super().__init__()
        conv = nn.ConvTranspose1d(cin, cout, kernel_size, stride=stride, padding=padding, output_padding=output_padding)
        v = conv.weight.detach()  # [cin, cout, k]; weight_norm dim=0 -> per INPUT channel
        self.weight_g = nn.Parameter(v.flatten(1).norm(dim=1).reshape(-1, 1, 1).clone())
        self.weight_v = nn.Parameter(v.clone())
        self.bias = nn.Parameter(conv.bias.detach().clone())
        self.stride, self.padding, self.output_padding = stride, padding, output_padding

    def forward(self, x):
        return F.conv_transpose1d(x, self._w(), self.bias, self.stride, self.padding, self.output_padding)


class Snake1d(nn.Module):
    def __init__(self, channels):
        super().__init__()
        self.alpha = nn.Parameter(torch.ones(1, channels, 1))

    def forward(self, x):
        return x + (self.alpha + 1e-9).reciprocal() * torch.sin(self.alpha * x).pow(2)


class NoiseBlock(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.linear = WNConv1d(dim, dim, 1, bias=False)
        self.source = "randn"  # "randn" | "off" | Tensor[B,1,T]

    def forward(self, x):
        B, _, T = x.shape
        if isinstance(self.source, torch.Tensor):
            n = self.source.to(x.dtype)
            if tuple(n.shape) != (B, 1, T):
                raise ValueError(f"injected noise has shape {tuple(n.shape)}, need {(B, 1, T)}")
        elif self.source == "off":
            n = torch.zeros((B, 1, T), dtype=x.dtype, device=x.device)
        else:
            n = torch.randn((B, 1, T), device=x.device, dtype=x.dtype)
        return x + n * self.linear(x)


class ResidualUnit(nn.Module):
    def __init__(self, dim, dilation):
        super().__init__()
        self.block = nn.Sequential(
            Snake1d(dim),
            WNConv1d(dim, dim, 7, padding=3 * dilation, dilation=dilation, groups=dim),
            Snake1d(dim),
            WNConv1d(dim, dim, 1),
        )

    def forward(self, x):
        return x + self.block(x)


class DecoderBlock(nn.Module):
    def __init__(self, cin, cout, stride):
        super().__init__()
        self.block = nn.Sequential(
            Snake1d(cin),
            WNConvTranspose1d(cin, cout, 2 * stride, stride, math.ceil(stride / 2), stride % 2),
            NoiseBlock(cout),
            ResidualUnit(cout, 1),
            ResidualUnit(cout, 3),
            ResidualUnit(cout, 9),
        )

    def forward(self, x):
        return self.block(x)


class Decoder(nn.Module):
    def __init__(self, latent, channels, rates):
        super().__init__()
        layers: List[nn.Module] = [
            WNConv1d(latent, latent, 7, padding=3, groups=latent),
            WNConv1d(latent, channels, 1),
        ]
        c = channels
        for s in rates:
            layers.append(DecoderBlock(c, c // 2, s))
            c //= 2
        layers += [Snake1d(c), WNConv1d(c, 1, 7, padding=3), nn.Tanh()]
        self.model = nn.Sequential(*layers)

    def forward(self, z):
        return self.model(z)


class EncoderBlock(nn.Module):
    """ResidualUnit d = 1, 3, 9 -> Snake -> strided WN Conv1d (k = 2s, stride s, pad ceil(s/2)); depthwise RUs."""

    def __init__(self, cin, cout, stride):
        super().__init__()
        self.block = nn.Sequential(
            ResidualUnit(cin, 1),
            ResidualUnit(cin, 3),
            ResidualUnit(cin, 9),
            Snake1d(cin),
            WNConv1d(cin, cout, 2 * stride, padding=math.ceil(stride / 2), stride=stride),
        )

    def forward(self, x):
        return self.block(x)


class Encoder(nn.Module):
    """Published SNAC encoder (snac_24khz: d_model 48, strides [2,4,8,8], depthwise, no attention):
    WN Conv1d 1->48 k7, 4 x EncoderBlock (channels double), WN depthwise Conv1d k7 on the 768 latent channels."""

    def __init__(self, d_model, rates):
        super().__init__()
        layers: List[nn.Module] = [WNConv1d(1, d_model, 7, padding=3)]
        for s in rates:
            layers.append(EncoderBlock(d_model, 2 * d_model, s))
            d_model *= 2
        layers.append(WNConv1d(d_model, d_model, 7, padding=3, groups=d_model))
        self.block = nn.Sequential(*layers)

    def forward(self, x):
        return self.block(x)


class VectorQuantize(nn.Module):
    def __init__(self, latent, size, dim, stride):
        super().__init__()
        self.in_proj = WNConv1d(latent, dim, 1)  # encode only
        self.out_proj = WNConv1d(dim, latent, 1)
        self.codebook = nn.Embedding(size, dim)
        self.stride = stride

    def encode(self, z):
        """Published VectorQuantize.forward: avg_pool -> in_proj -> nearest codebook entry on L2-normalised vectors ->
        out_proj -> repeat_interleave.  Returns (z_q [B,768,T], indices [B,T/stride])."""
        if self.stride > 1:
            z = F.avg_pool1d(z, self.stride, self.stride)
        z_e = self.in_proj(z)
        B, D, T = z_e.shape
        enc = F.normalize(z_e.transpose(1, 2).reshape(B * T, D))
        cb = F.normalize(self.codebook.weight)
        dist = enc.pow(2).sum(1, keepdim=True) - 2 * enc @ cb.t() + cb.pow(2).sum(1, keepdim=True).t()
        idx = (-dist).max(1)[1].reshape(B, T)
        z_q = self.out_proj(F.embedding(idx, self.codebook.weight).transpose(1, 2))
        if self.stride > 1:
            z_q = z_q.repeat_interleave(self.stride, dim=-1)
        return z_q, idx


class ResidualVectorQuantize(nn.Module):
    def __init__(self, latent, size, dim, strides):
        super().__init__()
        self.quantizers = nn.ModuleList(VectorQuantize(latent, size, dim, s) for s in strides)

    def encode(self, z):
        """Published ResidualVectorQuantize.forward: each level quantises what the previous ones left."""
        residual, codes = z, []
        for q in self.quantizers:
            z_q, idx = q.encode(residual)
            residual = residual - z_q
            codes.append(idx)
        return codes

    def from_codes(self, codes: Sequence[torch.Tensor]) -> torch.Tensor:
        z = 0.0
        for q, c in zip(self.quantizers, codes):
            e = F.embedding(c, q.codebook.weight).transpose(1, 2)  # raises IndexError for 4096 (Q1)
            z = z + q.out_proj(e).repeat_interleave(q.stride, dim=-1)
        return z


# --------------------------------------------------------------------------- model
_KEY_ALIASES = (
    (".parametrizations.weight.original0", ".weight_g"),
    (".parametrizations.weight.original1", ".weight_v"),
)


def normalise_keys(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    out = {}
    for k, v in sd.items():
        for new, old in _KEY_ALIASES:
            if k.endswith(new):
                k = k[: -len(new)] + old
        out[k] = v
    return out


# name -> state dict; lets tests make ``from_pretrained("hubertsiuzdak/snac_24khz")``
# (no network here) resolve to a known set of weights.
PRETRAINED: Dict[str, Dict[str, torch.Tensor]] = {}


class SNAC(nn.Module):
    """Decode-side restatement; the encoder (not on the reference's serving path) is attached when the
    state dict carries ``encoder.*`` keys."""

    def __init__(self, **cfg):
        super().__init__()
        c = dict(CONFIG_24KHZ)
        c.update({k: v for k, v in cfg.items() if k in c})
        self.config = c
        latent = c["latent_dim"] or c["encoder_dim"] * 2 ** len(c["encoder_rates"])
        self.latent_dim = latent
        self.hop_length = int(torch.prod(torch.tensor(c["decoder_rates"])))
        self.vq_strides = list(c["vq_strides"])
        self.quantizer = ResidualVectorQuantize(latent, c["codebook_size"], c["codebook_dim"], c["vq_strides"])
        self.decoder = Decoder(latent, c["decoder_dim"], c["decoder_rates"])
        self.encoder: Optional[Encoder] = None  # built by from_state_dict when the dict carries encoder.* keys

    # -- construction -------------------------------------------------------
    @classmethod
    def from_state_dict(cls, sd: Dict[str, torch.Tensor], **cfg) -> "SNAC":
        m = cls(**cfg)
        sd = normalise_keys(sd)
        if any(k.startswith("encoder.") for k in sd):
            m.encoder = Encoder(m.config["encoder_dim"], m.config["encoder_rates"])
        missing, unexpected = m.load_state_dict(sd, strict=False)
        missing = [k for k in missing if ".in_proj." not in k]
        if missing or unexpected:
            raise KeyError(f"state dict mismatch: missing={missing[:4]} unexpected={unexpected[:4]}")
        return m

    @classmethod
    def from_pretrained(cls, source: str, **kw) -> "SNAC":
        if source in PRETRAINED:
            return cls.from_state_dict(PRETRAINED[source])
        if os.path.isdir(source):
            with open(os.path.join(source, "config.json")) as f:
                cfg = json.load(f)
            sd = torch.load(os.path.join(source, "pytorch_model.bin"), map_location="cpu", weights_only=True)
            return cls.from_state_dict(sd, **cfg)
        # No network in this image: a hub id resolves to seeded default-init weights.
        gen_state = torch.random.get_rng_state()
        torch.manual_seed(0)
        m = cls()
        torch.random.set_rng_state(gen_state)
        return m

    # -- noise --------------------------------------------------------------
    def noise_blocks(self) -> List[NoiseBlock]:
        return [m for m in self.decoder.modules() if isinstance(m, NoiseBlock)]

    def set_noise(self, noise: Union[str, Sequence[torch.Tensor]]) -> None:
        blocks = self.noise_blocks()
        if isinstance(noise, str):
            for b in blocks:
                b.source = noise
        else:
            if len(noise) != len(blocks):
                raise ValueError("need one noise tensor per decoder block")
            for b, n in zip(blocks, noise):
                b.source = n

    # -- the path -----------------------------------------------------------
    def decode(self, codes: Sequence[torch.Tensor]) -> torch.Tensor:
        return self.decoder(self.quantizer.from_codes(codes))

    # -- N4: the encode direction (published SNAC.preprocess / SNAC.encode) ----
    def preprocess(self, audio: torch.Tensor) -> torch.Tensor:
        pad_to = self.hop_length * math.lcm(self.vq_strides[0], self.config["attn_window_size"] or 1)
        length = audio.shape[-1]
        return F.pad(audio, (0, math.ceil(length / pad_to) * pad_to - length))

    def encode_latent(self, audio: torch.Tensor) -> torch.Tensor:
        if self.encoder is None:
            raise RuntimeError("this oracle model was built without encoder weights")
        return self.encoder(self.preprocess(audio))

    def encode(self, audio: torch.Tensor) -> List[torch.Tensor]:
        """audio [B,1,T] -> codes [B,T'/4], [B,T'/2], [B,T'] with T' = padded length / 512."""
        return self.quantizer.encode(self.encode_latent(audio))


def noise_lengths(frames: int, rates=(8, 8, 4, 2)) -> List[int]:
    """Time length of the noise tensor of each decoder block for an F-frame decode."""
    t, out = 4 * frames, []
    for s in rates:
        t *= s
        out.append(t)
    return out


def make_noise(batch: int, frames: int, seed: int = 99) -> List[torch.Tensor]:
    """Mode-A noise of SURVEY 8(d): one CPU generator, blocks drawn in order."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    return [torch.randn((batch, 1, t), generator=g) for t in noise_lengths(frames)]


def pack_noise(noise: Sequence[torch.Tensor]) -> torch.Tensor:
    """[B,1,T_b] x4 -> [B, sum T_b] float32, the layout the C-ABI takes."""
    return torch.cat([n[:, 0, :] for n in noise], dim=1).contiguous()
