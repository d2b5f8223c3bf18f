"""``TTSAdapter`` over the B200 SNAC path, registrable with the reference's ``AdapterRegistry``.

Conforms to ``/root/reference/Morpheus_Client/orchestrator/adapter.py:13-60`` (``AudioChunk`` fields,
``pull(chunk_size)`` never returns more than ``chunk_size`` bytes and never blocks for the whole utterance,
``reset()`` after barge-in) and keeps the byte semantics the reference pins in ``tests/test_tts_adapter_chunking.py``
(every chunk but the last is exactly ``chunk_size`` bytes).  Token strings come from an injected async source (e.g. the
reference's ``remote_backend.generate_tokens_from_api``); the LLM side is out of scope.

How it is built (not the reference's pull-driven generator + ``bytearray``): the first ``pull`` starts one PUMP task
per adapter that runs ``speechpipe.tokens_decoder`` and writes decoded chunks into a :class:`~.ring.PcmRing`, ahead of
the consumer up to a high-water mark.  All pumps of the process decode through the shared
:class:`~.ticker.DecodeTicker`, so the windows of every live request form one GPU batch per tick, whatever the pull
cadence of the individual orchestrators is; ``pull`` itself only waits on the ring.

Registration (what a maintainer adds)::

    from project_morpheus_b200.adapter import register
    register(registry)            # registry.create("snac_b200", prompt=..., voice=..., seed=...)
"""
from __future__ import annotations

import asyncio
import inspect
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


class SnacB200Adapter:
    """Pull-based adapter: token strings -> sliding-window SNAC decode on the B200 -> PCM16 chunks."""

    name = "snac_b200"
    _default_source: TokenSource = staticmethod(_no_source)

    def __init__(self, prompt: str, voice: str = "tara", *, use_batching: bool = False, max_batch_chars: int = 1000,
                 seed: Optional[int] = None, token_source: Optional[TokenSource] = None,
                 decoder: Optional[Callable[..., AsyncIterator[bytes]]] = None, high_water: int = HIGH_WATER_BYTES,
                 **source_kwargs: Any) -> None:
        self.prompt, self.voice = prompt, voice
        self.use_batching, self.max_batch_chars = use_batching, max_batch_chars
        self.seed = seed  # NoiseBlock noise of this request: same seed + same tokens -> same bytes
        self._source = token_source
        self._source_kwargs = source_kwargs
        self._decoder = decoder
        self._high_water = max(1, int(high_water))
        self._ring = PcmRing()
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
        if decoder is None:
            from . import speechpipe  # CUDA path; no CPU fallback

            decoder = speechpipe.tokens_decoder
        if self.seed is not None and "stream_key" in inspect.signature(decoder).parameters:
            return decoder(tokens, stream_key=int(self.seed))
        return decoder(tokens)

    async def _run_pump(self, stream: AsyncIterator[bytes]) -> None:
        try:
            async for pcm in stream:
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
        return AudioChunk(pcm=pcm, duration_ms=len(pcm) / BYTES_PER_MS, eos=self._done and not len(self._ring))

    async def reset(self) -> None:
        """Barge-in: drop everything in flight; the next ``pull`` restarts the request."""
        pump, self._pump = self._pump, None
        if pump is not None:
            pump.cancel()
            try:
                await pump
            except BaseException:  # noqa: BLE001
                pass
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
