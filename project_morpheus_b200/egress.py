"""PCM egress: the reference's stitcher and RIFF framing over the native implementation.

* ``riff_header`` / ``wav_streamer``  <- ``Morpheus_Client/server.py:50-78``
* ``stitch_chunks``                   <- ``Morpheus_Client/orchestrator/stitcher.py:10-79`` (same signature, same
  ``AudioChunk`` objects, byte-identical PCM: the overlap-add runs in ``csrc/egress.cpp`` with numpy's float64
  arithmetic restated operation for operation)
* ``Stitcher``                        the same state machine as a plain push / flush object (many streams, no asyncio)
"""
from __future__ import annotations

import ctypes as C
from typing import AsyncGenerator, AsyncIterator, Optional, Tuple

import numpy as np

from . import _lib
from .adapter import AudioChunk

SAMPLE_RATE = 24000


def riff_header(sample_rate: int = SAMPLE_RATE) -> bytes:
    """Generic RIFF/WAVE header with unknown length (mono PCM16)."""
    buf = (C.c_uint8 * 44)()
    n = _lib.load().snacb_riff_header(int(sample_rate), buf)
    if n != 44:
        raise _lib.SnacbError(f"snacb_riff_header failed ({n})")
    return bytes(buf)


async def wav_streamer(pcm_iter, sample_rate: int = SAMPLE_RATE):
    """Wrap a PCM iterator with a WAV header for streaming."""
    yield riff_header(sample_rate)
    async for chunk in pcm_iter:
        yield chunk


class Stitcher:
    """Overlap-add joiner of one stream's chunks (the reference's ``stitch_chunks`` as push / flush)."""

    def __init__(self, sample_rate: int = SAMPLE_RATE, overlap_ms: float = 0.0):
        self._lib = _lib.load()
        self._h = C.c_void_p()
        rc = self._lib.snacb_stitch_create(C.byref(self._h), int(sample_rate), float(overlap_ms))
        if rc != _lib.OK:
            raise _lib.SnacbError(f"snacb_stitch_create failed ({rc})")
        self.sample_rate = int(sample_rate)
        self.overlap_samples = int(self._lib.snacb_stitch_overlap_samples(self._h))

    def close(self) -> None:
        if getattr(self, "_h", None) and self._h.value:
            self._lib.snacb_stitch_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = close

    def push(self, pcm: bytes, eos: bool = False) -> Tuple[Optional[bytes], bool]:
        """One chunk in -> (bytes the reference yields for it or None, eos flag of that yield)."""
        src = np.frombuffer(pcm, dtype="<i2")
        cap = len(src) + self.overlap_samples + 8
        out = np.empty(cap, dtype=np.int16)
        emitted, out_eos = C.c_int32(0), C.c_int32(0)
        n = self._lib.snacb_stitch_push(self._h, src.ctypes.data if len(src) else None, len(src), 1 if eos else 0,
                                        out.ctypes.data, cap, C.byref(emitted), C.byref(out_eos))
        if n < 0:
            raise _lib.SnacbError(f"snacb_stitch_push failed ({n})")
        if not emitted.value:
            return None, False
        return out[:n].tobytes(), bool(out_eos.value)

    def flush(self) -> Optional[bytes]:
        """The source ended without an eos chunk: the kept tail, or None when there is none."""
        cap = self.overlap_samples + 8
        out = np.empty(cap, dtype=np.int16)
        n = self._lib.snacb_stitch_flush(self._h, out.ctypes.data, cap)
        if n < 0:
            raise _lib.SnacbError(f"snacb_stitch_flush failed ({n})")
        return out[:n].tobytes() if n else None


class StitcherBank:
    """One stitcher per stream slot; ``push_tick`` joins a whole decode tick (rows of the PCM matrix) in one native call."""

    def __init__(self, n_streams: int, sample_rate: int = SAMPLE_RATE, overlap_ms: float = 0.0):
        self._lib = _lib.load()
        self._h = C.c_void_p()
        rc = self._lib.snacb_stitch_bank_create(C.byref(self._h), int(n_streams), int(sample_rate), float(overlap_ms))
        if rc != _lib.OK:
            raise _lib.SnacbError(f"snacb_stitch_bank_create failed ({rc})")
        self.overlap_samples = int(overlap_ms * sample_rate / 1000.0)

    def close(self) -> None:
        if getattr(self, "_h", None) and self._h.value:
            self._lib.snacb_stitch_bank_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = close

    def reset(self, slot: int) -> None:
        if self._lib.snacb_stitch_bank_reset(self._h, int(slot)) != _lib.OK:
            raise _lib.SnacbError("snacb_stitch_bank_reset failed")

    def push_tick(self, slots, pcm: np.ndarray, eos=None):
        """pcm int16 [n, len] (chunk i belongs to stream slots[i]) -> (out int16 [n, len + overlap], out_len int64 [n]
        with -1 where nothing is yielded, out_eos int32 [n])."""
        pcm = np.ascontiguousarray(pcm, dtype=np.int16)
        n, ln = pcm.shape
        sl = np.ascontiguousarray(slots, dtype=np.int32)
        ei = np.ascontiguousarray(eos, dtype=np.int32) if eos is not None else None
        out = np.empty((n, ln + self.overlap_samples), dtype=np.int16)
        out_len = np.empty(n, dtype=np.int64)
        out_eos = np.empty(n, dtype=np.int32)
        rc = self._lib.snacb_stitch_bank_push(self._h, n, sl.ctypes.data, pcm.ctypes.data, ln, ln,
                                              ei.ctypes.data if ei is not None else None, out.ctypes.data, out.shape[1],
                                              out_len.ctypes.data, out_eos.ctypes.data)
        if rc != _lib.OK:
            raise _lib.SnacbError(f"snacb_stitch_bank_push failed ({rc})")
        return out, out_len, out_eos


async def stitch_chunks(chunks: AsyncIterator[AudioChunk], *, sample_rate: int, overlap_ms: float = 0.0,
                        emit_markers: bool = False) -> AsyncGenerator[AudioChunk, None]:
    """Join ``chunks`` using overlap-add with optional marker propagation (reference signature and results)."""
    st = Stitcher(sample_rate, overlap_ms)
    try:
        ended = False
        async for chunk in chunks:
            data, eos = st.push(chunk.pcm, chunk.eos)
            if data is not None:
                yield AudioChunk(pcm=data, duration_ms=len(data) / 2 / sample_rate * 1000.0,
                                 markers=chunk.markers if emit_markers else None, eos=eos)
            if chunk.eos:
                ended = True
                break
        if not ended:
            tail = st.flush()
            if tail:
                yield AudioChunk(pcm=tail, duration_ms=len(tail) / 2 / sample_rate * 1000.0, markers=None, eos=True)
    finally:
        st.close()
