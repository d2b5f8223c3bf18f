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
import itertools
import logging
import os
from typing import AsyncIterator, List, Optional, Sequence

import numpy as np
import torch

from . import _lib
from . import ticker as ticker_mod
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
# Philox keys: a window decoded through tokens_decoder is keyed by (its stream's key, its index in the stream) -
# see ticker.window_key - so a stream's noise never depends on what other streams do.  Bare convert_to_audio calls
# carry no stream identity (the reference signature has none): they draw from this process-wide sequence.
_call_counter = itertools.count()
INT32_MIN, INT32_MAX = -(1 << 31), (1 << 31) - 1


def _decode_batch(tokens: np.ndarray, ntok: Optional[Sequence[int]], noise=None, keys: Optional[np.ndarray] = None):
    eng = model.engine
    n = tokens.shape[0]
    mode = noise if noise is not None else noise_mode
    if isinstance(mode, str) and mode == "philox":
        if keys is None:
            keys = np.asarray([ticker_mod.window_key(0, next(_call_counter)) for _ in range(n)], dtype=np.uint64)
    else:
        keys = None
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
    if status == _lib.WIN_NONFINITE:
        log.error("SNAC decode produced non-finite samples (fp16 operand range exceeded?); window withheld - "
                  "set SNACB_PRECISION=fp16x3 or fp32 for this checkpoint")
    return None


def _as_int32(ids: Sequence[int]) -> Optional[np.ndarray]:
    """Token ids as int32, or None when one does not fit (the reference's ``torch.tensor(frame, dtype=torch.int32)``
    raises for those, speechpipe.py:81; they are never wrapped into the valid code range here)."""
    a = np.asarray(ids, dtype=np.int64)
    if a.size and (int(a.min()) < INT32_MIN or int(a.max()) > INT32_MAX):
        return None
    return a.astype(np.int32)


def convert_to_audio(multiframe: Sequence[int], count: int) -> Optional[bytes]:
    """One window of token ids -> PCM16 bytes of samples [2048, 4096) of its decode (``count`` unused)."""
    if len(multiframe) < TOKENS_PER_FRAME:
        return None
    usable = (len(multiframe) // TOKENS_PER_FRAME) * TOKENS_PER_FRAME
    tokens = _as_int32(multiframe[:usable])
    if tokens is None:
        raise RuntimeError("value cannot be converted to type int32 without overflow")
    pcm, status = _decode_batch(tokens.reshape(1, usable), None)
    return _finish(pcm[0], int(status[0]))


def convert_to_audio_batch(windows: Sequence[Sequence[int]], noise=None, keys: Optional[np.ndarray] = None,
                           errors: str = "none") -> List[Optional[bytes]]:
    """All pending windows of a decode tick in one launch sequence; entry i is what
    ``convert_to_audio(windows[i], _)`` returns.  Where that call would RAISE (``IndexError`` for a 4096 code,
    ``RuntimeError`` for an id outside int32) entry i is ``None`` (``errors="none"``: one poisoned stream cannot fail
    the tick) or the exception instance itself (``errors="values"``: the ticker re-raises it in that stream only).
    ``keys``: one uint64 Philox key per window (default: the process-wide call sequence)."""
    n = len(windows)
    if n == 0:
        return []
    lens = [len(w) for w in windows]
    stride = max(TOKENS_PER_FRAME, max(lens))
    tokens = np.zeros((n, stride), dtype=np.int32)
    overflow = set()
    for i, w in enumerate(windows):
        if lens[i]:
            row = _as_int32(w)
            if row is None:
                if lens[i] >= TOKENS_PER_FRAME:  # shorter windows return None before the tensor is built (:69-70)
                    overflow.add(i)
                lens[i] = 0  # -> WIN_REJECTED
            else:
                tokens[i, : lens[i]] = row
    uniform = len(set(lens)) == 1 and lens[0] >= TOKENS_PER_FRAME
    pcm, status = _decode_batch(tokens, None if uniform else lens, noise=noise, keys=keys)
    out: List[Optional[bytes]] = []
    for i in range(n):
        st = int(status[i])
        if i in overflow:
            out.append(RuntimeError("value cannot be converted to type int32 without overflow") if errors == "values" else None)
        elif st == _lib.WIN_CODE4096:
            out.append(IndexError("index out of range in self") if errors == "values" else None)
        else:
            out.append(_finish(pcm[i], st))
    return out


# ----------------------------------------------------------------------------- GPU egress rings (N3)
_ring = None


def get_ring():
    """The process-wide :class:`~project_morpheus_b200.egress.GpuPcmRing` (created on first use): one pinned ring per live
    stream that the decode tick fills on the GPU.  ``SNACB_RING_SLOTS`` (1024) streams x ``SNACB_RING_SAMPLES`` (32768 =
    1.37 s) samples; ``SNACB_RING_OVERLAP_MS`` (0) crossfades consecutive windows of a stream like the reference's
    stitcher does with a non-zero overlap."""
    global _ring
    if _ring is None:
        from .egress import GpuPcmRing

        _ring = GpuPcmRing(int(os.environ.get("SNACB_RING_SLOTS", "1024")), int(os.environ.get("SNACB_RING_SAMPLES", "32768")),
                           overlap_ms=float(os.environ.get("SNACB_RING_OVERLAP_MS", "0")),
                           device=torch.cuda.current_device() if snac_device == "cuda" else 0)
    return _ring


def convert_to_ring_batch(windows: Sequence[Sequence[int]], slots: Sequence[int], keys: Optional[np.ndarray] = None,
                          errors: str = "none", ring=None) -> list:
    """``convert_to_audio_batch`` with the PCM of window i appended to ring slot ``slots[i]`` on the GPU instead of being
    returned: entry i is the number of PCM bytes the window added to its ring (4096 without crossfade, 0 for a single-frame
    window) where ``convert_to_audio`` returns bytes, and ``None`` / an exception exactly where ``convert_to_audio_batch`` has them."""
    n = len(windows)
    if n == 0:
        return []
    ring = ring if ring is not None else get_ring()
    lens = [len(w) for w in windows]
    stride = max(TOKENS_PER_FRAME, max(lens))
    tokens = np.zeros((n, stride), dtype=np.int32)
    overflow = set()
    for i, w in enumerate(windows):
        if lens[i]:
            row = _as_int32(w)
            if row is None:
                if lens[i] >= TOKENS_PER_FRAME:
                    overflow.add(i)
                lens[i] = 0
            else:
                tokens[i, : lens[i]] = row
    uniform = len(set(lens)) == 1 and lens[0] >= TOKENS_PER_FRAME
    mode = noise_mode if noise_mode in ("philox", "off") else "philox"
    if mode == "philox" and keys is None:
        keys = np.asarray([ticker_mod.window_key(0, next(_call_counter)) for _ in range(n)], dtype=np.uint64)
    eng = model.engine

    def run():
        return eng.decode_windows_to_ring(ring, tokens, slots, ntok=None if uniform else lens, noise=mode, seed=model.noise_seed,
                                          keys=keys if mode == "philox" else None)

    if cuda_stream is not None:
        with torch.cuda.stream(cuda_stream):
            status, emitted = run()
    else:
        status, emitted = run()
    out: list = []
    for i in range(n):
        st = int(status[i])
        if i in overflow:
            out.append(RuntimeError("value cannot be converted to type int32 without overflow") if errors == "values" else None)
        elif st == _lib.WIN_CODE4096:
            out.append(IndexError("index out of range in self") if errors == "values" else None)
        elif st == _lib.WIN_OK and emitted[i] < 0:
            # no room in that slot: the window was decoded to nowhere, an error of that stream alone (adapters never get
            # here: their pump stops at a high-water mark)
            out.append(BufferError("PCM ring slot is full: read it before decoding more") if errors == "values" else None)
        elif st == _lib.WIN_OK:
            out.append(2 * int(emitted[i]))  # 4096, or what the crossfade released when an overlap is configured
        elif st == _lib.WIN_EMPTY:
            out.append(0)
        else:
            out.append(_finish(None, st))  # None (+ the NONFINITE log line)
    return out


def _tick(windows, keys, slots=None):
    """One decode tick of the shared ticker: ring-bound windows (slot >= 0) and plain ones, each kind in one engine call."""
    if slots is None:
        return convert_to_audio_batch(windows, keys=keys, errors="values")
    to_ring = [i for i, s in enumerate(slots) if s >= 0]
    plain = [i for i, s in enumerate(slots) if s < 0]
    out: list = [None] * len(windows)
    if to_ring:
        res = convert_to_ring_batch([windows[i] for i in to_ring], [slots[i] for i in to_ring],
                                    keys=None if keys is None else keys[to_ring], errors="values")
        for i, r in zip(to_ring, res):
            out[i] = r
    if plain:
        res = convert_to_audio_batch([windows[i] for i in plain], keys=None if keys is None else keys[plain], errors="values")
        for i, r in zip(plain, res):
            out[i] = r
    return out


# ----------------------------------------------------------------------------- the shared decode ticker
USE_TICKER = os.environ.get("SNACB_TICKER", "1") != "0"
_ticker: Optional["ticker_mod.DecodeTicker"] = None


def get_ticker() -> "ticker_mod.DecodeTicker":
    """The process-wide ticker every ``tokens_decoder`` coroutine decodes through (created on first use)."""
    global _ticker
    if _ticker is None:
        _ticker = ticker_mod.DecodeTicker(lambda *a: _tick(*a))  # late-bound: tests swap the batch functions
    return _ticker


from .tokens import (  # noqa: E402,F401  (re-exported: same names as the reference module)
    CUSTOM_TOKEN_PREFIX, FIRST_WINDOW, LONG_WINDOW, MAX_CACHE_SIZE, SHORT_WINDOW, WindowPlanner, token_id_cache,
    turn_token_into_id,
)


async def tokens_decoder(token_gen: AsyncIterator[str], *, stream_key: Optional[int] = None, ticker=None,
                         ring_slot: Optional[int] = None):
    """Token strings in, PCM chunks out; same windows, order and chunk sizes as the reference.

    Additive keywords: ``stream_key`` seeds this stream's NoiseBlock noise (window w of the stream uses the Philox key
    ``window_key(stream_key, w)``; default: a fresh key per stream); ``ticker`` overrides the shared
    :class:`~project_morpheus_b200.ticker.DecodeTicker` (``False`` = decode each window on its own, like the
    reference).  With the ticker, the windows of all concurrently running ``tokens_decoder`` coroutines of the event
    loop go to the GPU as one batch per tick.  ``ring_slot``: the PCM is appended to that slot of ``get_ring()`` by the
    GPU and the generator yields byte COUNTS (4096 / 0) in place of the bytes (``SnacB200Adapter(gpu_ring=True)``)."""
    plan = WindowPlanner()
    key = ticker_mod.fresh_stream_key() if stream_key is None else int(stream_key)
    tk = get_ticker() if (ticker is None and USE_TICKER) else (ticker or None)
    n_windows = 0

    async def decode(window):
        nonlocal n_windows
        wkey = ticker_mod.window_key(key, n_windows)
        n_windows += 1
        if tk is not None:  # raises what convert_to_audio would raise, in this stream only
            return await (tk.decode(list(window), wkey) if ring_slot is None else tk.decode(list(window), wkey, ring_slot))
        if ring_slot is None:
            out = convert_to_audio_batch([window], keys=np.asarray([wkey], dtype=np.uint64), errors="values")[0]
        else:
            out = convert_to_ring_batch([window], [ring_slot], keys=np.asarray([wkey], dtype=np.uint64), errors="values")[0]
        if isinstance(out, BaseException):
            raise out
        return out

    async for token_sim in token_gen:
        window = plan.push(token_sim)
        if window is None:
            continue
        audio_samples = await decode(window)
        plan.result(audio_samples)
        if audio_samples is not None:
            yield audio_samples
    window = plan.flush()
    if window is not None:
        audio_samples = await decode(window)
        if audio_samples is not None:
            yield audio_samples
    if ring_slot is not None and get_ring().overlap_samples > 0:
        before = get_ring().available(ring_slot)
        get_ring().flush(ring_slot)  # the crossfade tail kept from the last window
        yield get_ring().available(ring_slot) - before


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
