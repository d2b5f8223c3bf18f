"""Shared helpers for the parity tests (test infrastructure; may import ``oracle``)."""
from __future__ import annotations

import json
import os
from typing import List, Sequence

import numpy as np
import torch

from oracle import snac_ref, speechpipe_ref as sp

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "speechpipe_golden.json")

# north_star tolerance for the waveform: max-abs <= 2e-3 AND SNR >= 60 dB in fp32 (or <= 2 LSB in int16)
TOL_MAX_ABS = 2e-3
TOL_SNR_DB = 60.0
# the fp32 CUDA-core recipe only re-associates sums: far tighter
TOL_FP32_MAX_ABS = 2e-5


def load_golden() -> dict:
    with open(GOLDEN) as f:
        return json.load(f)


def snr_db(ref: np.ndarray, got: np.ndarray) -> float:
    ref = np.asarray(ref, dtype=np.float64)
    err = np.asarray(got, dtype=np.float64) - ref
    den = float(np.sum(err * err))
    if den == 0.0:
        return float("inf")
    return 10.0 * np.log10(float(np.sum(ref * ref)) / den)


def windows_tokens(n: int, frames: int, base_stream: int = 0) -> np.ndarray:
    """[n, 7*frames] int32 token ids, SURVEY 8(d) recipe (PCG64(1234+s), codes U{1..4095})."""
    return np.stack([sp.synth_codes(base_stream + i, frames) for i in range(n)]).astype(np.int32)


def oracle_decode_windows(model: snac_ref.SNAC, tokens: np.ndarray, noise: Sequence[torch.Tensor] | str) -> np.ndarray:
    """Full fp32 decode of every window through the oracle: [n, 2048*F] float32."""
    n = tokens.shape[0]
    lv = [sp.split_levels(list(map(int, row))) for row in tokens]
    codes = [torch.from_numpy(np.stack([l[k] for l in lv]).astype(np.int64)) for k in range(3)]
    model.set_noise(noise)
    with torch.no_grad():
        y = model.decode(codes)
    model.set_noise("off")
    return y[:, 0, :].numpy()


def pcm_trunc(x: np.ndarray) -> np.ndarray:
    return (np.asarray(x, dtype=np.float32) * np.float32(32767)).astype(np.int16)
