"""Drop-in replacement for ``Morpheus_Client/tts_engine/speechpipe.py`` on B200 kernels.

Same public names, signatures and return conventions as the reference module
(``/root/reference/Morpheus_Client/tts_engine/speechpipe.py``):

* ``convert_to_audio(multiframe, count) -> bytes | None``   (reference ``:64-137``)
* ``turn_token_into_id(token_string, index) -> int | None`` (reference ``:146-189``)
* ``tokens_decoder(token_gen)``  async generator of bytes   (reference ``:191-293``)
* ``tokens_decoder_sync(syn_token_gen)``                    (reference ``:295-337``)
* module attributes ``model``, ``snac_device``, ``cuda_stream``, ``CUSTOM_TOKEN_PREFIX``,
  ``token_id_cache``, ``MAX_CACHE_SIZE``.

so ``remote_backend.py:30`` and the upstream ``engine_class.py:10`` import it unchanged.
Additive: ``convert_to_audio_batch(windows)`` decodes a whole tick in ONE engine call
(the reference serialises one B=1 decode per stream per tick).

Return conventions kept: ``None`` = skipped window (fewer than 7 tokens, or a code outside
``[0, 4096]``), ``b''`` = a single-frame window (its ``[2048:4096)`` slice is empty), else 4096
bytes of little-endian int16 PCM.  A code equal to 4096 raises ``IndexError`` exactly like the
reference's embedding lookup does.  There is no CPU fallback: without CUDA the decode raises.
"""
from __future__ import annotations

import asyncio
import logging
import os
from typing import AsyncIterator, List, Optional, Sequence

import numpy as np
import torch

from . import _lib
from .snac import SNAC

log = logging.getLogger("project_morpheus_b200.speechpipe")

# speechpipe.py:38-43 - ORPHEUS_SNAC_PATH overrides the hub id; the path string goes to from_pretrained.
snac_path = os.environ.get("ORPHEUS_SNAC_PATH")
model_source = snac_path if snac_path else "hubertsiuzdak/snac_24khz"
model = SNAC.from_pretrained(model_source).eval()

# speechpipe.py:46-49 - device pick.  On a CUDA host the engine (weights + workspace) is built here,
# at import, like the reference; on a CPU-only host import succeeds and decode raises.
snac_device = "cuda" if torch.cuda.is_available() else "cpu"
model = model.to(snac_device)

# speechpipe.py:57-59 - one side stream for all requests.
cuda_stream = torch.cuda.Stream() if snac_device == "cuda" else None

TOKENS_PER_FRAME = 7
PCM_BYTES = 4096

# Noise for NoiseBlock: "philox" (fresh in-kernel noise per window, like the reference's randn),
# "off", or inject explicit tensors through convert_to_audio_batch(noise=...).
noise_mode = os.environ.get("SNACB_NOISE", "philox")
_window_counter = 0


def _decode_batch(tokens: np.ndarray, ntok: Optional[Sequence[int]], noise=None):
    global _window_counter
    eng = model.engine
    n = tokens.shape[0]
    mode = noise if noise is not None else noise_mode
    keys = None
    if isinstance(mode, str) and mode == "philox":
        keys = np.arange(_window_counter, _window_counter + n, dtype=np.uint64)
        _window_counter += n
    if cuda_stream is not None:
        with torch.cuda.stream(cuda_stream):
            return eng.decode_windows(tokens, ntok=ntok, noise=mode, seed=model.noise_seed, keys=keys)
    return eng.decode_windows(tokens, ntok=ntok, noise=mode, seed=model.noise_seed, keys=keys)


def _finish(pcm_row: np.ndarray, status: int) -> Optional[bytes]:
    if status == _lib.WIN_OK:
        return pcm_row.tobytes()
    if status == _lib.WIN_EMPTY:
        return b""
    if status == _lib.WIN_CODE4096:
        raise IndexError("index out of range in self")
    return None


def convert_to_audio(multiframe: Sequence[int], count: int) -> Optional[bytes]:
    """One window of token ids -> PCM16 bytes of samples [2048, 4096) of its decode (``count`` unused)."""
    if len(multiframe) < TOKENS_PER_FRAME:
        return None
    usable = (len(multiframe) // TOKENS_PER_FRAME) * TOKENS_PER_FRAME
    tokens = np.asarray(multiframe[:usable], dtype=np.int64).astype(np.int32).reshape(1, usable)
    pcm, status = _decode_batch(tokens, None)
    return _finish(pcm[0], int(status[0]))


def convert_to_audio_batch(windows: Sequence[Sequence[int]], noise=None) -> List[Optional[bytes]]:
    """All pending windows of a decode tick in one launch sequence; entry i is what
    ``convert_to_audio(windows[i], _)`` returns (``IndexError`` for a 4096 code is reported as
    ``None`` here so one poisoned stream cannot fail the tick)."""
    n = len(windows)
    if n == 0:
        return []
    lens = [len(w) for w in windows]
    stride = max(TOKENS_PER_FRAME, max(lens))
    if min(lens) == stride:  # uniform tick (the common case): one vectorised conversion
        tokens = np.asarray(windows, dtype=np.int64).astype(np.int32).reshape(n, stride)
    else:
        tokens = np.zeros((n, stride), dtype=np.int32)
        for i, w in enumerate(windows):
            if lens[i]:
                tokens[i, : lens[i]] = np.asarray(w, dtype=np.int64).astype(np.int32)
    uniform = len(set(lens)) == 1 and lens[0] >= TOKENS_PER_FRAME
    pcm, status = _decode_batch(tokens, None if uniform else lens, noise=noise)
    out: List[Optional[bytes]] = []
    for i in range(n):
        st = int(status[i])
        out.append(None if st == _lib.WIN_CODE4096 else _finish(pcm[i], st))
    return out


from .tokens import (  # noqa: E402,F401  (re-exported: same names as the reference module)
    CUSTOM_TOKEN_PREFIX, FIRST_WINDOW, LONG_WINDOW, MAX_CACHE_SIZE, SHORT_WINDOW, WindowPlanner, token_id_cache,
    turn_token_into_id,
)


async def tokens_decoder(token_gen: AsyncIterator[str]):
    """Token strings in, PCM chunks out; same windows, order and chunk sizes as the reference."""
    plan = WindowPlanner()
    async for token_sim in token_gen:
        window = plan.push(token_sim)
        if window is None:
            continue
        audio_samples = convert_to_audio(window, plan.count)
        plan.result(audio_samples)
        if audio_samples is not None:
            yield audio_samples
    window = plan.flush()
    if window is not None:
        audio_samples = convert_to_audio(window, plan.count)
        if audio_samples is not None:
            yield audio_samples


async def tokens_decoder_sync(syn_token_gen):
    """Queue-decoupled variant: drops empty chunks and releases audio in groups of five."""
    audio_queue: asyncio.Queue = asyncio.Queue(maxsize=32 if snac_device == "cuda" else 8)

    async def producer():
        try:
            async for chunk in tokens_decoder(syn_token_gen):
                if chunk:
                    await audio_queue.put(chunk)
        except Exception:  # noqa: BLE001 - the reference swallows producer errors and ends the stream
            log.exception("error in audio producer")
        finally:
            await audio_queue.put(None)

    task = asyncio.create_task(producer())
    held: List[bytes] = []
    while True:
        chunk = await audio_queue.get()
        if chunk is None:
            break
        held.append(chunk)
        if len(held) >= 5:
            for c in held:
                yield c
            held = []
    for c in held:
        yield c
    await task
