"""Process-wide decode ticker: the tick batcher behind the reference's seam.

The reference's orchestrator pulls ONE adapter per request (``/root/reference/Morpheus_Client/orchestrator/core.py:89-117``)
and the server builds one adapter per HTTP call (``server.py:144-156``); each adapter drives its own ``tokens_decoder``
(``tts_engine/speechpipe.py:191-293``), which calls ``convert_to_audio`` once per window - one B = 1 decode per stream
per tick.  Nothing in that control flow is changed here.  What changes is what a decode call does while it waits:

* every ``tokens_decoder`` coroutine of the process hands its window to the shared :class:`DecodeTicker` and awaits a
  future;
* one ticker task per event loop collects whatever windows are pending after the loop has run every other ready
  coroutine once, and submits them as ONE ``convert_to_audio_batch`` call (one launch sequence on the GPU) in a worker
  thread (``asyncio.to_thread``, the reference's own idiom for blocking work, ``llama_local.py:79``);
* windows that arrive while a tick is on the GPU form the next tick.

So N concurrent requests cost one batched decode per tick instead of N serial ones, with the orchestrator, the server
and the adapter protocol untouched.  Per-stream results are independent of batching: a window's bytes depend only on
its tokens and its own Philox key ``(stream key, per-stream window index)``.
"""
from __future__ import annotations

import asyncio
import itertools
import threading
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np

BatchDecode = Callable[..., List[Optional[bytes]]]  # (windows, keys[, slots]) -> one entry per window

_SPLITMIX = (0x9E3779B97F4A7C15, 0xBF58476D1CE4E5B9, 0x94D049BB133111EB)
_MASK = (1 << 64) - 1


def mix64(x: int) -> int:
    """splitmix64 finaliser: spreads (stream key, window index) over the 64-bit Philox key space."""
    x = (x + _SPLITMIX[0]) & _MASK
    x = ((x ^ (x >> 30)) * _SPLITMIX[1]) & _MASK
    x = ((x ^ (x >> 27)) * _SPLITMIX[2]) & _MASK
    return x ^ (x >> 31)


def window_key(stream_key: int, window_index: int) -> int:
    """Noise key of window ``window_index`` of the stream ``stream_key``: a function of the stream alone, never of how
    the calls of concurrent streams interleave (replayable per stream, identical under any partitioning)."""
    return mix64((mix64(stream_key & _MASK) + window_index) & _MASK)


_stream_counter = itertools.count(1)
_stream_lock = threading.Lock()


def fresh_stream_key() -> int:
    """Key of a stream that was not given a seed: the creation rank of the stream in this process."""
    with _stream_lock:
        return next(_stream_counter)


class DecodeTicker:
    """Coalesces the decode calls of every coroutine of one event loop into ticks."""

    def __init__(self, batch_decode: BatchDecode, max_batch: int = 4096, in_thread: bool = True,
                 settle_turns: int = 1) -> None:
        self._decode = batch_decode
        self.max_batch = int(max_batch)
        self.in_thread = in_thread
        self.settle_turns = max(0, int(settle_turns))
        self._pending: List[Tuple[Sequence[int], int, asyncio.Future, Optional[int]]] = []
        self._wake: Optional[asyncio.Event] = None
        self._idle: Optional[asyncio.Event] = None  # set whenever no tick is on the GPU
        self._inflight: set = set()                 # ring slots the running tick writes to
        self._task: Optional[asyncio.Task] = None
        self._loop: Optional[asyncio.AbstractEventLoop] = None
        self.ticks = 0
        self.windows = 0
        self.max_tick = 0

    # ------------------------------------------------------------------ client side
    async def decode(self, window: Sequence[int], key: int = 0, slot: Optional[int] = None):
        """What ``convert_to_audio(window, _)`` returns, decoded together with every other pending window.

        ``slot``: the window's PCM goes to that slot of the GPU egress ring instead of coming back as bytes (the result is
        then the byte count).  A caller that is cancelled while its window is queued or on the GPU must ``await
        forget(slot)`` before the slot is reset or re-used."""
        loop = asyncio.get_running_loop()
        if self._loop is not loop:  # first use on this loop (or the previous loop is gone): (re)start the ticker task
            self._bind(loop)
        fut: asyncio.Future = loop.create_future()
        self._pending.append((window, key, fut, slot))
        self._wake.set()
        return await fut

    async def forget(self, slot: int) -> None:
        """Drop the queued windows of ``slot`` and wait until no tick that writes to it is on the GPU."""
        keep = []
        for p in self._pending:
            if p[3] == slot:
                if not p[2].done():
                    p[2].cancel()
            else:
                keep.append(p)
        self._pending = keep
        while slot in self._inflight and self._idle is not None and asyncio.get_running_loop() is self._loop:
            await self._idle.wait()

    def _bind(self, loop: asyncio.AbstractEventLoop) -> None:
        for _, _, fut, _ in self._pending:  # requests of a dead loop can never be answered
            if not fut.done():
                fut.cancel()
        self._pending = []
        self._loop = loop
        self._wake = asyncio.Event()
        self._idle = asyncio.Event()
        self._idle.set()
        self._inflight = set()
        self._task = loop.create_task(self._run(), name="snacb-decode-ticker")

    # ------------------------------------------------------------------ ticker task
    async def _run(self) -> None:
        wake = self._wake
        while True:
            await wake.wait()
            wake.clear()
            # let every coroutine that is runnable right now reach its own decode call before the tick is cut
            for _ in range(self.settle_turns):
                await asyncio.sleep(0)
            while self._pending:
                batch, self._pending = self._pending[: self.max_batch], self._pending[self.max_batch:]
                windows = [b[0] for b in batch]
                keys = np.asarray([b[1] & _MASK for b in batch], dtype=np.uint64)
                args = (windows, keys)
                if any(b[3] is not None for b in batch):  # ring-bound windows: the batch function also gets the slots
                    args = (windows, keys, [-1 if b[3] is None else int(b[3]) for b in batch])
                    self._inflight = {b[3] for b in batch if b[3] is not None}
                    self._idle.clear()
                try:
                    if self.in_thread:
                        out = await asyncio.to_thread(self._decode, *args)
                    else:
                        out = self._decode(*args)
                    err = None
                except BaseException as e:  # noqa: BLE001 - handed to every waiter of the tick
                    out, err = None, e
                if self._inflight:
                    self._inflight = set()
                    self._idle.set()
                self.ticks += 1
                self.windows += len(batch)
                self.max_tick = max(self.max_tick, len(batch))
                for i, (_, _, fut, _) in enumerate(batch):
                    if fut.done():
                        continue
                    if err is not None:
                        fut.set_exception(err if not isinstance(err, asyncio.CancelledError) else RuntimeError("tick cancelled"))
                    elif isinstance(out[i], BaseException):
                        fut.set_exception(out[i])  # what convert_to_audio raises for this window; other streams go on
                    else:
                        fut.set_result(out[i])
                if err is not None and isinstance(err, (KeyboardInterrupt, SystemExit, asyncio.CancelledError)):
                    raise err

    async def aclose(self) -> None:
        task, self._task, self._loop = self._task, None, None
        if task is not None:
            task.cancel()
            try:
                await task
            except BaseException:  # noqa: BLE001
                pass

    def stats(self) -> Dict[str, float]:
        return {"ticks": self.ticks, "windows": self.windows, "max_tick": self.max_tick,
                "windows_per_tick": (self.windows / self.ticks) if self.ticks else 0.0}
