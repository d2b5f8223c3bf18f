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


class GpuPcmRing:
    """Pinned per-stream PCM rings the decode tick writes on the GPU (``csrc/egress_ring.cu``).

    Replaces, for all streams of a tick at once, the reference's per-request chain stitcher -> ring buffer -> ``pull``
    re-chunking (``orchestrator/stitcher.py:10-79``, ``orchestrator/ring_buffer.py:27-83``,
    ``tts_engine/llama_local.py:120-150``): consecutive chunks of a slot are crossfaded by a kernel right after the decoder
    tail (byte-identical to the reference's float64 overlap-add; ``overlap_ms = 0`` is plain concatenation, the server's
    default) and written through the rings' device mapping; ``read`` is a cursor move and one memcpy out of pinned
    memory.  Needs a CUDA device: there is no host fallback."""

    def __init__(self, n_slots: int, ring_samples: int = 1 << 15, sample_rate: int = SAMPLE_RATE, overlap_ms: float = 0.0,
                 device: int = 0):
        self._lib = _lib.load()
        self._h = C.c_void_p()
        rc = self._lib.snacb_egress_create(C.byref(self._h), int(device), int(n_slots), int(ring_samples), int(sample_rate),
                                           float(overlap_ms))
        if rc != _lib.OK:
            raise _lib.SnacbError(f"snacb_egress_create failed ({rc}): needs a CUDA device, ring_samples a multiple of 8 "
                                  f">= 4096 and 2 * overlap + 2048 <= ring_samples")
        self.n_slots, self.ring_samples, self.device = int(n_slots), int(ring_samples), int(device)
        self.overlap_samples = int(self._lib.snacb_egress_overlap_samples(self._h))
        self._free = list(range(self.n_slots - 1, -1, -1))
        # zero-call reads: views of the pinned rings and of both cursor arrays
        wp, rp = C.c_void_p(), C.c_void_p()
        if self._lib.snacb_egress_cursors(self._h, C.byref(wp), C.byref(rp)) != _lib.OK:
            raise _lib.SnacbError("snacb_egress_cursors failed")
        self._wpos = np.ctypeslib.as_array((C.c_int64 * self.n_slots).from_address(wp.value))
        self._rpos = np.ctypeslib.as_array((C.c_int64 * self.n_slots).from_address(rp.value))
        base = self._lib.snacb_egress_ring_base(self._h, 0)
        self._ring_mv = memoryview((C.c_char * (2 * self.n_slots * self.ring_samples)).from_address(base)).cast("B")

    @property
    def handle(self) -> C.c_void_p:
        return self._h

    def close(self) -> None:
        if getattr(self, "_h", None) and self._h.value:
            self._wpos = self._rpos = self._ring_mv = None  # views into memory the library is about to free
            self._lib.snacb_egress_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = close

    def _check(self, rc: int, what: str) -> None:
        if rc < 0:
            raise _lib.SnacbError(f"{what} failed ({rc}): {self._lib.snacb_egress_last_error(self._h).decode(errors='replace')}")

    # slot bookkeeping (one slot per live stream)
    def acquire(self) -> int:
        if not self._free:
            raise _lib.SnacbError(f"GpuPcmRing: all {self.n_slots} slots are in use")
        return self._free.pop()

    def release(self, slot: int, stream: Optional[int] = None) -> None:
        self.reset(slot, stream)
        self._free.append(int(slot))

    def push_device(self, slots, d_pcm_ptr: int, pcm_stride: int, length: int, d_status_ptr: Optional[int] = None, eos=None,
                    stream: Optional[int] = None) -> None:
        """Asynchronous: rows of a DEVICE int16 matrix join the rings (``sync`` before reading)."""
        sl = np.ascontiguousarray(slots, dtype=np.int32)
        eo = np.ascontiguousarray(eos, dtype=np.int32) if eos is not None else None
        self._check(self._lib.snacb_egress_push_device(self._h, len(sl), sl.ctypes.data, d_pcm_ptr, int(pcm_stride), int(length),
                                                       d_status_ptr, eo.ctypes.data if eo is not None else None, stream),
                    "snacb_egress_push_device")

    def sync(self, stream: Optional[int] = None) -> None:
        self._check(self._lib.snacb_egress_sync(self._h, stream), "snacb_egress_sync")

    def available(self, slot: int) -> int:
        """Unread bytes of the slot."""
        return 2 * (int(self._wpos[slot]) - int(self._rpos[slot]))

    def room(self, slot: int) -> int:
        """Samples a tick may still add to the slot."""
        n = int(self._lib.snacb_egress_room(self._h, int(slot)))
        self._check(n, "snacb_egress_room")
        return n

    def read(self, slot: int, nbytes: int) -> bytes:
        """Up to ``nbytes`` (rounded down to whole samples) of the slot's unread PCM: one copy out of the pinned ring, no
        native call (the read cursor lives in memory shared with the library, see ``snacb_egress_cursors``)."""
        r = int(self._rpos[slot])
        n = min(max(0, int(nbytes)) // 2, int(self._wpos[slot]) - r)
        if n <= 0:
            return b""
        R = self.ring_samples
        off = r % R
        b0 = 2 * (slot * R + off)
        if off + n <= R:
            data = self._ring_mv[b0: b0 + 2 * n].tobytes()
        else:
            first = R - off
            data = self._ring_mv[b0: b0 + 2 * first].tobytes() + self._ring_mv[2 * slot * R: 2 * (slot * R + n - first)].tobytes()
        self._rpos[slot] = r + n
        return data

    def flush(self, slot: int, stream: Optional[int] = None) -> None:
        """End of the stream without an eos chunk: the kept crossfade tail is emitted."""
        self._check(self._lib.snacb_egress_flush(self._h, int(slot), stream), "snacb_egress_flush")

    def reset(self, slot: int, stream: Optional[int] = None) -> None:
        self._check(self._lib.snacb_egress_reset(self._h, int(slot), stream), "snacb_egress_reset")


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
