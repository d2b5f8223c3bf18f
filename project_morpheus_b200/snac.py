"""Drop-in for the third-party ``snac`` module as the reference uses it.

The reference does exactly four things with ``snac``
(``/root/reference/Morpheus_Client/tts_engine/speechpipe.py:1,43,49,118``):
``from snac import SNAC``; ``SNAC.from_pretrained(src).eval()``; ``.to(device)``;
``model.decode(codes)``.  Installing this module as ``sys.modules['snac']`` lets the
reference file run unmodified on the B200 kernels (same injection the reference's own
``tests/test_speechpipe_snac_path.py:7-22,39`` performs with a dummy).

There is no CPU path: ``.to('cpu')`` is accepted (the reference calls it on CPU-only
hosts at import time) but ``decode`` then raises.
"""
from __future__ import annotations

import os
from typing import Dict, Optional, Sequence

import torch

from . import weights as _weights

DEFAULT_REPO = "hubertsiuzdak/snac_24khz"


def _resolve_state_dict(source: str) -> Dict[str, torch.Tensor]:
    """``source``: local dir (config.json + pytorch_model.bin), ``random:<seed>[:w1]``, or a hub id."""
    if os.path.isdir(source):
        return _weights.load_checkpoint(source)
    spec = source if source.startswith("random:") else os.environ.get("SNACB_RANDOM_INIT")
    if spec is not None:
        parts = spec.replace("random:", "").split(":")
        seed = int(parts[0]) if parts[0] else 0
        variant = parts[1] if len(parts) > 1 else "default"
        return _weights.random_state_dict(seed, variant)
    try:  # offline hub cache, if the checkpoint was ever downloaded
        from huggingface_hub import hf_hub_download  # type: ignore

        cfg = hf_hub_download(source, "config.json", local_files_only=True)
        return _weights.load_checkpoint(os.path.dirname(cfg))
    except Exception as exc:  # noqa: BLE001
        raise FileNotFoundError(
            f"SNAC weights {source!r} not found locally and there is no network: set ORPHEUS_SNAC_PATH to a "
            "directory holding config.json + pytorch_model.bin, or SNACB_RANDOM_INIT=<seed> for seeded random weights"
        ) from exc


class SNAC:
    """The four-method surface of ``snac.SNAC`` the reference touches, over ``SnacEngine``."""

    sampling_rate = 24000
    hop_length = 512
    vq_strides = [4, 2, 1]

    def __init__(self, source: str = DEFAULT_REPO, state_dict: Optional[Dict[str, torch.Tensor]] = None):
        self.source = source
        self._state_dict = state_dict
        self._engine = None
        self.device = torch.device("cpu")
        self.training = False
        # NoiseBlock draws fresh randn per call in the reference; production uses in-kernel Philox.
        self.noise = os.environ.get("SNACB_NOISE", "philox")
        seed_env = os.environ.get("SNACB_NOISE_SEED")
        self.noise_seed = int(seed_env) if seed_env is not None else int.from_bytes(os.urandom(8), "little")
        self.precision = os.environ.get("SNACB_PRECISION", "fp16")
        self._calls = 0

    @classmethod
    def from_pretrained(cls, source: str, **kwargs) -> "SNAC":
        return cls(source)

    @classmethod
    def from_state_dict(cls, state_dict: Dict[str, torch.Tensor], precision: Optional[str] = None,
                        noise: Optional[str] = None) -> "SNAC":
        m = cls("<state_dict>", state_dict)
        if precision is not None:
            m.precision = precision
        if noise is not None:
            m.noise = noise
        return m

    def eval(self) -> "SNAC":
        self.training = False
        return self

    def to(self, device) -> "SNAC":
        dev = torch.device(device)
        if dev.type == "cuda":
            self.device = torch.device("cuda", dev.index if dev.index is not None else 0)
            self._ensure_engine()
        else:
            self.device = dev
        return self

    # ------------------------------------------------------------------
    def _ensure_engine(self):
        if self._engine is None:
            if self.device.type != "cuda":
                raise RuntimeError(
                    f"SNAC decode on device {self.device} is not available: this build has no CPU fallback "
                    "(move the model with .to('cuda'))"
                )
            from .engine import SnacEngine

            sd = self._state_dict if self._state_dict is not None else _resolve_state_dict(self.source)
            self._engine = SnacEngine(sd, device=self.device.index or 0, precision=self.precision)
        return self._engine

    @property
    def engine(self):
        return self._ensure_engine()

    def encode(self, audio_data: torch.Tensor):
        """``snac.SNAC.encode``: float audio ``[B,1,T]`` -> three code tensors (needs a checkpoint that
        carries the encoder weights, e.g. the published ``pytorch_model.bin``)."""
        return self._ensure_engine().encode(audio_data)

    def decode(self, codes: Sequence[torch.Tensor]) -> torch.Tensor:
        """``[B,F],[B,2F],[B,4F]`` integer codes -> float32 ``[B,1,2048F]`` on the model's device."""
        eng = self._ensure_engine()
        for c in codes:
            if c.numel() and (int(c.min()) < 0 or int(c.max()) >= _weights.CODEBOOK_SIZE):
                raise IndexError("index out of range in self")  # what F.embedding raises in the reference
        self._calls += 1
        noise = self.noise
        return eng.decode_codes(codes, noise=noise, seed=self.noise_seed + self._calls)
