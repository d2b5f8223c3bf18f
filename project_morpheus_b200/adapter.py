"""``TTSAdapter`` over the B200 SNAC path, registrable with the reference's ``AdapterRegistry``.

Conforms to ``/root/reference/Morpheus_Client/orchestrator/adapter.py:13-60`` (``AudioChunk`` fields,
``pull(chunk_size)`` never returns more than ``chunk_size`` bytes and never blocks for the whole utterance,
``reset()`` after barge-in) and keeps the byte semantics the reference pins in ``tests/test_tts_adapter_chunking.py``
(every chunk but the last is exactly ``chunk_size`` bytes).  Token strings come from an injected async source (e.g. the
reference's ``remote_backend.generate_tokens_from_api``); the LLM side is out of scope.

How it is built (not the reference's pull-driven generator + ``bytearray``): the first ``pull`` starts one PUMP task
per adapter that runs ``speechpipe.tokens_decoder`` ahead of the consumer up to a high-water mark.  All pumps of the
process decode through the shared :class:`~.ticker.DecodeTicker`, so the windows of every live request form one GPU
batch per tick, whatever the pull cadence of the individual orchestrators is; ``pull`` itself only waits on the ring.
Where the decoded audio waits for ``pull``:

* ``gpu_ring=True`` (default on a CUDA host, ``SNACB_GPU_RING=0`` turns it off): in a slot of the process-wide
  :class:`~.egress.GpuPcmRing` - pinned memory the decode tick writes ON THE GPU (``csrc/egress_ring.cu``), so the
  tick's PCM never comes back as a matrix that Python slices; ``pull`` is one native read of the slot;
* otherwise in a host-side :class:`~.ring.PcmRing` fed with the ``bytes`` chunks of the decoder.

Registration (what a maintainer adds)::

    from project_morpheus_b200.adapter import register
    register(registry)            # registry.create("snac_b200", prompt=..., voice=..., seed=...)
"""
from __future__ import annotations

import asyncio
import inspect
import os
from dataclasses import dataclass
from typing import Any, AsyncIterator, Callable, Dict, Optional

from .ring import PcmRing

SAMPLE_RATE = 24000
BYTES_PER_MS = SAMPLE_RATE * 2 / 1000.0
HIGH_WATER_BYTES = 8 * 4096  # decoded audio the pump may run ahead of the consumer (8 windows = 0.68 s)

try:  # the reference's own dataclass when the package is importable, else an identical stand-in
    from Morpheus_Client.orchestrator.adapter import AudioChunk  # type: ignore
except Exception:  # noqa: BLE001

    @dataclass
    class AudioChunk:  # same fields and defaults as orchestrator/adapter.py:13-35
        pcm: bytes
        duration_ms: float
        markers: Optional[object] = None
        eos: bool = False


TokenSource = Callable[..., AsyncIterator[str]]


def _no_source(**_: Any) -> AsyncIterator[str]:
    raise RuntimeError(
        "SnacB200Adapter needs a token_source (async iterator of '<custom_token_N>' strings); "
        "pass token_source=... or call adapter.configure_token_source(fn)"
    )


class _GpuSlot:
    """One slot of the GPU egress ring behind the small interface the adapter uses of :class:`~.ring.PcmRing`."""

    def __init__(self, ring: Any) -> None:
        self._ring, self.slot = ring, ring.acquire()
        self._carry = b""  # the ring hands out whole samples; an odd pull leaves one byte here
        self._avail = 0    # bytes the decoder reported as written and nobody has read yet (no native call per len())

    def __len__(self) -> int:
        return self._avail + len(self._carry)

    def wrote(self, nbytes: int) -> None:
        self._avail += nbytes

    def read(self, n: int) -> bytes:
        need = n - len(self._carry)
        if n <= 0 or need <= 0:
            out, self._carry = self._carry[: max(0, n)], self._carry[max(0, n):]
            return out
        got = self._ring.read(self.slot, need + (need & 1))
        self._avail -= len(got)
        if not self._carry and len(got) <= n:
            return got
        data = self._carry + got
        self._carry = data[n:]
        return data[:n]

    def clear(self) -> None:
        self._carry, self._avail = b"", 0
        self._ring.reset(self.slot)

    def close(self) -> None:
        if self.slot >= 0:
            self._ring.release(self.slot)
            self.slot = -1


class SnacB200Adapter:
    """Pull-based adapter: token strings -> sliding-window SNAC decode on the B200 -> PCM16 chunks."""

    name = "snac_b200"
    _default_source: TokenSource = staticmethod(_no_source)

    def __init__(self, prompt: str, voice: str = "tara", *, use_batching: bool = False, max_batch_chars: int = 1000,
                 seed: Optional[int] = None, token_source: Optional[TokenSource] = None,
                 decoder: Optional[Callable[..., AsyncIterator[bytes]]] = None, high_water: int = HIGH_WATER_BYTES,
                 gpu_ring: Optional[bool] = None, **source_kwargs: Any) -> None:
        self.prompt, self.voice = prompt, voice
        self.use_batching, self.max_batch_chars = use_batching, max_batch_chars
        self.seed = seed  # NoiseBlock noise of this request: same seed + same tokens -> same bytes
        self._source = token_source
        self._source_kwargs = source_kwargs
        self._decoder = decoder
        self._high_water = max(1, int(high_water))
        # None: decided at the first pull (GPU ring when the product decoder runs on a CUDA host and SNACB_GPU_RING != 0)
        self._gpu_ring = gpu_ring
        self._ring: Any = PcmRing()
        self._pump: Optional[asyncio.Task] = None
        self._changed: Optional[asyncio.Event] = None  # ring gained data / drained below the mark / stream ended
        self._done = False
        self._error: Optional[BaseException] = None

    @classmethod
    def configure_token_source(cls, fn: TokenSource) -> None:
        cls._default_source = staticmethod(fn)

    # ------------------------------------------------------------------ producer side
    def _open_stream(self) -> AsyncIterator[bytes]:
        src = self._source or type(self)._default_source
        tokens = src(prompt=self.prompt, voice=self.voice, **self._source_kwargs)
        decoder = self._decoder
        kwargs: Dict[str, Any] = {}
        if decoder is None:
            from . import speechpipe  # CUDA path; no CPU fallback

            decoder = speechpipe.tokens_decoder
            want = self._gpu_ring
            if want is None:
                want = speechpipe.snac_device == "cuda" and os.environ.get("SNACB_GPU_RING", "1") != "0"
            if want and not isinstance(self._ring, _GpuSlot):
                try:
                    self._ring = _GpuSlot(speechpipe.get_ring())
                    self._high_water = min(self._high_water, 2 * (speechpipe.get_ring().ring_samples - 4096))
                except speechpipe._lib.SnacbError as e:
                    if self._gpu_ring:  # asked for explicitly
                        raise
                    # more live requests than ring slots (SNACB_RING_SLOTS): this request keeps its PCM in host bytes
                    speechpipe.log.warning("GPU PCM ring unavailable for this request (%s): using the host ring", e)
            if isinstance(self._ring, _GpuSlot):
                kwargs["ring_slot"] = self._ring.slot
        if self.seed is not None and "stream_key" in inspect.signature(decoder).parameters:
            kwargs["stream_key"] = int(self.seed)
        return decoder(tokens, **kwargs)

    async def _run_pump(self, stream: AsyncIterator[bytes]) -> None:
        try:
            async for pcm in stream:
                if isinstance(pcm, int):  # that many bytes are already in the GPU ring slot
                    self._ring.wrote(pcm)
                else:
                    self._ring.write(pcm)
                self._changed.set()
                while len(self._ring) >= self._high_water:  # back-pressure: wait for the consumer
                    self._changed.clear()
                    if len(self._ring) < self._high_water:
                        break
                    await self._changed.wait()
        except asyncio.CancelledError:
            raise
        except BaseException as e:  # noqa: BLE001 - surfaced by the pull that finds the ring empty
            self._error = e
        finally:
            self._done = True
            self._changed.set()
            aclose = getattr(stream, "aclose", None)
            if aclose is not None:
                try:
                    await aclose()
                except BaseException:  # noqa: BLE001
                    pass
            if isinstance(self._ring, _GpuSlot) and self._decoder is None:
                # cancelled mid-decode (barge-in, dropped request): the slot may only be reset or handed to another
                # stream once no tick that writes to it is queued or on the GPU
                try:
                    from . import speechpipe

                    await speechpipe.get_ticker().forget(self._ring.slot)
                except BaseException:  # noqa: BLE001
                    pass

    def _start(self) -> None:
        if self._pump is None and not self._done:
            stream = self._open_stream()  # raises here (in pull) when no token source is configured
            self._changed = asyncio.Event()
            self._pump = asyncio.get_running_loop().create_task(self._run_pump(stream), name="snacb-adapter-pump")

    # ------------------------------------------------------------------ TTSAdapter protocol
    async def pull(self, chunk_size: int) -> AudioChunk:
        """Up to ``chunk_size`` PCM bytes: exactly ``chunk_size`` while the stream lasts, the remainder at its end."""
        want = max(0, int(chunk_size))
        self._start()
        while len(self._ring) < want and not self._done:
            self._changed.clear()
            if len(self._ring) >= want or self._done:
                break
            await self._changed.wait()
        if self._error is not None and not len(self._ring):
            err, self._error = self._error, None
            raise err
        pcm = self._ring.read(want)
        if self._changed is not None:
            self._changed.set()  # the pump may be waiting for room
        eos = self._done and not len(self._ring)
        if eos:
            self._release()
        return AudioChunk(pcm=pcm, duration_ms=len(pcm) / BYTES_PER_MS, eos=eos)

    def _release(self) -> None:
        """Give the GPU ring slot back (stream fully delivered, reset, or the adapter is dropped)."""
        if isinstance(self._ring, _GpuSlot):
            self._ring.close()
            self._ring = PcmRing()

    def __del__(self) -> None:
        try:
            if self._pump is None or self._pump.done():
                self._release()
        except Exception:  # noqa: BLE001 - interpreter shutdown
            pass

    async def reset(self) -> None:
        """Barge-in: drop everything in flight; the next ``pull`` restarts the request."""
        pump, self._pump = self._pump, None
        if pump is not None:
            pump.cancel()
            try:
                await pump
            except BaseException:  # noqa: BLE001
                pass
        self._release()  # the pump has ended: nothing of this stream is on the GPU any more
        self._ring.clear()
        self._done = False
        self._error = None
        self._changed = None


def describe() -> Dict[str, Any]:
    """Capability record in the shape the reference registry publishes (adapter_registry.py:48-60)."""
    return {
        "name": "snac_b200", "streaming": True, "unit": "bytes", "granularity": [8, 12, 16, 24, 32, 48, 64],
        "voices": ["tara", "leah", "jess", "leo", "dan", "mia", "zac", "zoe"],
        "supports_barge_in": True, "supports_seed": True, "stateful_context": False,
    }


def voice_mapper(voice: Any) -> Dict[str, Any]:
    return {"voice": getattr(voice, "name", None) or getattr(voice, "voice", None) or str(voice)}


def register(registry: Any, name: str = "snac_b200") -> None:
    """``registry`` is the reference's ``tts_engine.adapter_registry.registry`` (adapter_registry.py:76-83)."""
    registry.register(name, SnacB200Adapter, describe, voice_mapper)
