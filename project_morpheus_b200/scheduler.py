"""Tick scheduler: one engine launch sequence per decode tick for ALL live streams.

The reference drives one B=1 ``convert_to_audio`` per stream per 7 accepted tokens
(``/root/reference/Morpheus_Client/tts_engine/speechpipe.py:191-293``); many streams therefore
serialise hundreds of tiny decodes.  Here every stream keeps the reference's per-stream window
state machine (``tokens.WindowPlanner``: same windows, same order, same end-of-stream flush), the
windows that became ready since the last tick are decoded in ONE batched call
(``convert_to_audio_batch`` -> ``snacb_decode_windows_host``), and each stream receives exactly the
byte chunks ``tokens_decoder`` would have yielded for it.  Streams can be inserted and evicted
between ticks (barge-in, mid-stream swap); windows are ragged (7 / 28 / 49 tokens).

The decode function is injectable so the control flow is testable without a GPU; the default is the
CUDA path (there is no CPU fallback).
"""
from __future__ import annotations

from collections import deque
from typing import Callable, Deque, Dict, Hashable, Iterable, List, Optional, Sequence, Tuple

from .tokens import WindowPlanner

DecodeBatch = Callable[[Sequence[Sequence[int]]], List[Optional[bytes]]]


def _default_decode() -> DecodeBatch:
    from . import speechpipe  # imports the engine; fails loudly without CUDA at decode time

    return speechpipe.convert_to_audio_batch


class _Stream:
    __slots__ = ("plan", "pending", "out", "finished", "flushed")

    def __init__(self) -> None:
        self.plan = WindowPlanner()
        self.pending: Deque[Tuple[List[int], bool]] = deque()  # (window, is_first_chunk_probe)
        self.out: Deque[bytes] = deque()
        self.finished = False   # producer signalled end of stream
        self.flushed = False    # end-of-stream window already queued


class TickScheduler:
    """Batches the sliding windows of many token streams into one decode per tick."""

    def __init__(self, decode_batch: Optional[DecodeBatch] = None, max_windows_per_tick: int = 4096):
        self._decode = decode_batch
        self.max_windows_per_tick = int(max_windows_per_tick)
        self._streams: Dict[Hashable, _Stream] = {}
        self.ticks = 0
        self.windows_decoded = 0

    # ---------------------------------------------------------------- stream lifecycle
    def add_stream(self, sid: Hashable) -> None:
        if sid in self._streams:
            raise KeyError(f"stream {sid!r} already exists")
        self._streams[sid] = _Stream()

    def evict(self, sid: Hashable) -> None:
        """Drop a stream and everything queued for it (barge-in / swap): nothing of it is decoded again."""
        self._streams.pop(sid, None)

    def reset_stream(self, sid: Hashable) -> None:
        self._streams[sid] = _Stream()

    def __contains__(self, sid: Hashable) -> bool:
        return sid in self._streams

    @property
    def live_streams(self) -> int:
        return len(self._streams)

    # ---------------------------------------------------------------- token ingress
    def push(self, sid: Hashable, token_string: str) -> None:
        """Feed one token string of stream ``sid`` (same acceptance rules as ``tokens_decoder``)."""
        st = self._streams[sid]
        if st.finished:
            raise RuntimeError(f"stream {sid!r} already finished")
        # The first-chunk latch of the reference flips only after the 7-token decode returned non-None,
        # i.e. it depends on a decode result; until that result is known later tokens must wait.
        st.pending.append((token_string, None))  # type: ignore[arg-type]

    def push_many(self, sid: Hashable, token_strings: Iterable[str]) -> None:
        for t in token_strings:
            self.push(sid, t)

    def finish(self, sid: Hashable) -> None:
        self._streams[sid].finished = True

    # ---------------------------------------------------------------- the tick
    def _next_window(self, st: _Stream) -> Optional[List[int]]:
        """Advance the stream's planner over its queued token strings up to the next window."""
        while st.pending:
            token_string, _ = st.pending.popleft()
            win = st.plan.push(token_string)
            if win is not None:
                return list(win)
        if st.finished and not st.flushed:
            st.flushed = True
            return st.plan.flush()
        return None

    def tick(self) -> int:
        """Decode, in one batch, the next ready window of every stream; returns #windows decoded.

        One window per stream per tick keeps per-stream ordering trivially identical to the
        reference (a stream's next window may depend on the outcome of its previous one)."""
        if self._decode is None:
            self._decode = _default_decode()
        batch: List[List[int]] = []
        owners: List[_Stream] = []
        for st in self._streams.values():
            if len(batch) >= self.max_windows_per_tick:
                break
            win = self._next_window(st)
            if win is not None:
                batch.append(win)
                owners.append(st)
        if not batch:
            return 0
        results = self._decode(batch)
        if len(results) != len(batch):
            raise RuntimeError("decode_batch returned a different number of results")
        for st, audio in zip(owners, results):
            st.plan.result(audio)
            if audio is not None:
                st.out.append(audio)
        self.ticks += 1
        self.windows_decoded += len(batch)
        return len(batch)

    def drain(self) -> int:
        """Tick until no stream has a ready window; returns total windows decoded."""
        total = 0
        while True:
            n = self.tick()
            if n == 0:
                return total
            total += n

    # ---------------------------------------------------------------- PCM egress
    def pop_audio(self, sid: Hashable) -> List[bytes]:
        st = self._streams[sid]
        out = list(st.out)
        st.out.clear()
        return out

    def done(self, sid: Hashable) -> bool:
        st = self._streams[sid]
        return st.finished and st.flushed and not st.pending
