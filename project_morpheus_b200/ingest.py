"""Native token ingress + tick scheduler.

``NativeTickScheduler`` has the interface and the per-stream results of ``scheduler.TickScheduler`` (itself the
batched form of the reference ``tokens_decoder``, ``Morpheus_Client/tts_engine/speechpipe.py:191-293``), but the
per-token work - parsing ``<custom_token_N>`` (``speechpipe.py:146-189``), the drop-and-shift slot rule, the
sliding-window state of every stream - runs in C++ behind ``snacb_ingest_*`` (``csrc/ingest.cpp``): token strings of
all streams go in as one blob per call, a tick returns the ready windows as the int32 matrix the engine takes as is.
At 1024+ streams the Python planner, not the GPU, bounds a tick; this removes it.

There is no CPU decode: the default decode is the CUDA engine behind ``speechpipe``; tests inject one.
"""
from __future__ import annotations

import ctypes as C
from collections import deque
from typing import Callable, Deque, Dict, Hashable, Iterable, List, Optional, Sequence, Tuple

import numpy as np

from . import _lib

STRIDE = 49
DecodeArrays = Callable[[np.ndarray, np.ndarray], Tuple[np.ndarray, np.ndarray]]  # (tokens [n,49], ntok [n]) -> (pcm [n,2048] i16, status [n])
DecodeBatch = Callable[[Sequence[Sequence[int]]], List[Optional[bytes]]]


class NativeIngest:
    """Thin owner of a ``snacb_ingest`` handle: ``n_streams`` slots."""

    def __init__(self, n_streams: int):
        self._lib = _lib.load()
        self._h = C.c_void_p()
        rc = self._lib.snacb_ingest_create(C.byref(self._h), int(n_streams))
        if rc != _lib.OK:
            raise _lib.SnacbError(f"snacb_ingest_create failed ({rc})")
        self.n_streams = int(n_streams)
        self._tok = np.zeros((max(self.n_streams, 1), STRIDE), dtype=np.int32)
        self._ntok = np.zeros((max(self.n_streams, 1),), dtype=np.int32)
        self._owner = np.zeros((max(self.n_streams, 1),), dtype=np.int32)

    def close(self) -> None:
        if getattr(self, "_h", None) and self._h.value:
            self._lib.snacb_ingest_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = close

    def _check(self, rc: int, what: str) -> None:
        if rc < 0:
            raise _lib.SnacbError(f"{what} failed ({rc})")

    def reset(self, slot: int) -> None:
        self._check(self._lib.snacb_ingest_reset(self._h, int(slot)), "snacb_ingest_reset")

    def push(self, slots: Sequence[int], strings: Sequence[str]) -> None:
        """Token strings (one generator item each) for the given slots, one native call."""
        n = len(strings)
        if n == 0:
            return
        offs = np.zeros(n + 1, dtype=np.int64)
        np.cumsum(np.fromiter(map(len, strings), dtype=np.int64, count=n), out=offs[1:])
        blob = "".join(strings).encode("utf-8", "surrogatepass")  # one encode for the whole batch
        if len(blob) != int(offs[n]):  # some string is not ASCII: byte lengths differ from character counts
            enc = [t.encode("utf-8", "surrogatepass") for t in strings]
            np.cumsum(np.fromiter(map(len, enc), dtype=np.int64, count=n), out=offs[1:])
            blob = b"".join(enc)
        sl = np.ascontiguousarray(np.asarray(slots, dtype=np.int32))
        self._check(self._lib.snacb_ingest_push(self._h, n, sl.ctypes.data, blob, offs.ctypes.data), "snacb_ingest_push")

    def finish(self, slot: int) -> None:
        self._check(self._lib.snacb_ingest_finish(self._h, int(slot)), "snacb_ingest_finish")

    def tick(self, max_windows: Optional[int] = None) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        """(tokens [n,49] int32, ntok [n], slot [n]) views of the next ready window of every stream."""
        cap = self._tok.shape[0] if max_windows is None else min(int(max_windows), self._tok.shape[0])
        n = self._lib.snacb_ingest_tick(self._h, cap, self._tok.ctypes.data, STRIDE, self._ntok.ctypes.data,
                                        self._owner.ctypes.data)
        self._check(n, "snacb_ingest_tick")
        return self._tok[:n], self._ntok[:n], self._owner[:n]

    def result(self, slots: np.ndarray, status: np.ndarray) -> None:
        sl = np.ascontiguousarray(slots, dtype=np.int32)
        st = np.ascontiguousarray(status, dtype=np.int32)
        self._check(self._lib.snacb_ingest_result(self._h, len(sl), sl.ctypes.data, st.ctypes.data), "snacb_ingest_result")

    def done(self, slot: int) -> bool:
        return self._lib.snacb_ingest_done(self._h, int(slot)) == 1

    def stats(self) -> Dict[str, int]:
        return {k: int(self._lib.snacb_ingest_stat(self._h, i)) for i, k in enumerate(("accepted", "rejected", "windows"))}


def _default_decode_arrays() -> DecodeArrays:
    from . import speechpipe  # imports the engine; fails loudly without CUDA at decode time

    def run(tokens: np.ndarray, ntok: np.ndarray):
        return decode_arrays_with(speechpipe._decode_batch, tokens, ntok)

    return run


def decode_arrays_with(decode, tokens: np.ndarray, ntok: np.ndarray):
    """Call ``decode(tokens, ntok_or_None)``: a uniform tick (every window the same length - the steady state of
    many concurrent streams) takes the engine's uniform path, a ragged one passes the per-window lengths."""
    n0 = int(ntok[0])
    if n0 >= 7 and bool((ntok == n0).all()):
        return decode(np.ascontiguousarray(tokens[:, :n0]), None)
    return decode(tokens, ntok.tolist())


def _wrap_list_decode(decode: DecodeBatch) -> DecodeArrays:
    """Adapter for an injected ``convert_to_audio_batch``-shaped callable (tests, other back ends)."""

    def run(tokens: np.ndarray, ntok: np.ndarray):
        wins = [tokens[i, : int(ntok[i])].tolist() for i in range(len(ntok))]
        res = decode(wins)
        pcm = np.zeros((len(wins), 2048), dtype=np.int16)
        st = np.zeros((len(wins),), dtype=np.int32)
        for i, r in enumerate(res):
            if r is None:
                st[i] = _lib.WIN_REJECTED
            elif len(r) == 0:
                st[i] = _lib.WIN_EMPTY
            else:
                pcm[i] = np.frombuffer(r, dtype="<i2")
        return pcm, st

    return run


class NativeTickScheduler:
    """``TickScheduler`` with the token parsing and window planning of all streams in native code."""

    def __init__(self, decode_batch: Optional[DecodeBatch] = None, max_streams: int = 4096,
                 max_windows_per_tick: int = 4096, decode_arrays: Optional[DecodeArrays] = None,
                 engine=None, noise: str = "philox", seed: int = 0):
        """``engine`` (a ``SnacEngine``) switches to pipelined ticks: ``tick()`` submits the windows that are ready now
        (``submit_windows``) and returns the audio of the PREVIOUS tick (``wait_windows``), so token parsing, window
        planning, PCM distribution and the device -> host copy of one tick run under the kernels of the next.  Per
        stream the windows, their order and their bytes are unchanged; audio arrives one tick later."""
        self._eng = engine
        self._noise, self._seed, self._key0 = noise, int(seed), 0
        self._inflight = None
        self._ing = NativeIngest(max_streams)
        self._decode: Optional[DecodeArrays] = decode_arrays or (_wrap_list_decode(decode_batch) if decode_batch else None)
        self.max_windows_per_tick = int(max_windows_per_tick)
        self._slot: Dict[Hashable, int] = {}
        self._free: List[int] = list(range(max_streams - 1, -1, -1))
        self._sid_of: List[Optional[Hashable]] = [None] * max_streams
        self._gen: List[int] = [0] * max_streams  # bumped on reset / evict: in-flight windows of an old tenant are dropped
        self._out: Dict[Hashable, Deque[bytes]] = {}
        self._finished: Dict[Hashable, bool] = {}
        self._q_slots: List[int] = []
        self._q_text: List[str] = []
        self.ticks = 0
        self.windows_decoded = 0

    # ---------------------------------------------------------------- stream lifecycle
    def add_stream(self, sid: Hashable) -> None:
        if sid in self._slot:
            raise KeyError(f"stream {sid!r} already exists")
        if not self._free:
            raise RuntimeError("no free stream slot")
        slot = self._free.pop()
        self._ing.reset(slot)
        self._gen[slot] += 1
        self._slot[sid] = slot
        self._sid_of[slot] = sid
        self._out[sid] = deque()
        self._finished[sid] = False

    def evict(self, sid: Hashable) -> None:
        slot = self._slot.pop(sid, None)
        if slot is None:
            return
        self._flush_queue()
        self._ing.reset(slot)
        self._gen[slot] += 1
        self._sid_of[slot] = None
        self._free.append(slot)
        self._out.pop(sid, None)
        self._finished.pop(sid, None)

    def reset_stream(self, sid: Hashable) -> None:
        self._flush_queue()
        self._ing.reset(self._slot[sid])
        self._gen[self._slot[sid]] += 1
        self._out[sid] = deque()
        self._finished[sid] = False

    def __contains__(self, sid: Hashable) -> bool:
        return sid in self._slot

    @property
    def live_streams(self) -> int:
        return len(self._slot)

    # ---------------------------------------------------------------- token ingress
    def push(self, sid: Hashable, token_string: str) -> None:
        if self._finished[sid]:
            raise RuntimeError(f"stream {sid!r} already finished")
        self._q_slots.append(self._slot[sid])
        self._q_text.append(token_string)

    def push_many(self, sid: Hashable, token_strings: Iterable[str]) -> None:
        if self._finished[sid]:
            raise RuntimeError(f"stream {sid!r} already finished")
        slot = self._slot[sid]
        n0 = len(self._q_text)
        self._q_text.extend(token_strings)
        self._q_slots.extend([slot] * (len(self._q_text) - n0))

    def _flush_queue(self) -> None:
        if self._q_text:
            self._ing.push(self._q_slots, self._q_text)
            self._q_slots, self._q_text = [], []

    def finish(self, sid: Hashable) -> None:
        self._flush_queue()
        self._finished[sid] = True
        self._ing.finish(self._slot[sid])

    # ---------------------------------------------------------------- the tick
    def _deliver(self, pcm: np.ndarray, status: np.ndarray, owner: np.ndarray) -> None:
        sid_of, out = self._sid_of, self._out
        for i in np.nonzero(status == _lib.WIN_OK)[0].tolist():
            q = out.get(sid_of[owner[i]])
            if q is not None:  # the stream may have been evicted while its window was in flight
                q.append(pcm[i].tobytes())
        for i in np.nonzero(status == _lib.WIN_EMPTY)[0].tolist():
            q = out.get(sid_of[owner[i]])
            if q is not None:
                q.append(b"")

    def _tick_pipelined(self) -> int:
        self._flush_queue()
        tok, ntok, owner = self._ing.tick(self.max_windows_per_tick)
        n = len(ntok)
        nxt = None
        if n:
            keys = np.arange(self._key0, self._key0 + n, dtype=np.uint64)
            self._key0 += n
            n0 = int(ntok[0])
            if n0 >= 7 and bool((ntok == n0).all()):
                ticket = self._eng.submit_windows(tok[:, :n0], noise=self._noise, seed=self._seed, keys=keys)
            else:
                ticket = self._eng.submit_windows(tok, ntok=ntok, noise=self._noise, seed=self._seed, keys=keys)
            gen = [self._gen[s] for s in owner.tolist()]
            nxt = (ticket, owner.copy(), gen)
        done = 0
        if self._inflight is not None:
            ticket, own, gen = self._inflight
            pcm, status = self._eng.wait_windows(ticket)
            status = np.asarray(status)
            live = np.fromiter((self._gen[s] == g for s, g in zip(own.tolist(), gen)), dtype=bool, count=len(gen))
            if live.all():
                self._ing.result(own, status)
                self._deliver(pcm, status, own)
            elif live.any():  # windows of slots that were reset / re-used meanwhile are dropped
                self._ing.result(own[live], status[live])
                self._deliver(pcm[live], status[live], own[live])
            done = len(own)
            self.ticks += 1
            self.windows_decoded += done
        self._inflight = nxt
        return done if done else (-1 if nxt is not None else 0)  # -1: nothing delivered yet, but a tick is in flight

    def tick(self) -> int:
        if self._eng is not None:
            r = self._tick_pipelined()
            return 0 if r < 0 else r
        if self._decode is None:
            self._decode = _default_decode_arrays()
        self._flush_queue()
        tok, ntok, owner = self._ing.tick(self.max_windows_per_tick)
        n = len(ntok)
        if n == 0:
            return 0
        pcm, status = self._decode(tok, ntok)
        status = np.asarray(status)
        self._ing.result(owner, status)
        # a 4096 code (IndexError in the reference's embedding lookup) yields nothing for that stream, like
        # convert_to_audio_batch reports it: one poisoned stream cannot fail the tick
        ok = np.nonzero(status == _lib.WIN_OK)[0]
        sid_of, out = self._sid_of, self._out
        for i in ok.tolist():
            out[sid_of[owner[i]]].append(pcm[i].tobytes())
        for i in np.nonzero(status == _lib.WIN_EMPTY)[0].tolist():
            out[sid_of[owner[i]]].append(b"")
        self.ticks += 1
        self.windows_decoded += n
        return n

    def drain(self) -> int:
        total = 0
        while True:
            if self._eng is not None:
                r = self._tick_pipelined()
                if r == 0:
                    return total
                total += max(r, 0)
                continue
            n = self.tick()
            if n == 0:
                return total
            total += n

    # ---------------------------------------------------------------- PCM egress
    def pop_audio(self, sid: Hashable) -> List[bytes]:
        q = self._out[sid]
        res = list(q)
        q.clear()
        return res

    def done(self, sid: Hashable) -> bool:
        self._flush_queue()
        return self._ing.done(self._slot[sid])

    def close(self) -> None:
        self._ing.close()
