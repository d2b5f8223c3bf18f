"""``TTSAdapter`` over the B200 SNAC path, registrable with the reference's ``AdapterRegistry``.

Conforms to ``/root/reference/Morpheus_Client/orchestrator/adapter.py:13-60`` (``AudioChunk`` fields,
``pull(chunk_size)`` never returns more than ``chunk_size`` bytes and never blocks for the whole
utterance, ``reset()`` after barge-in) and mirrors the byte re-chunking of the reference's local
adapter (``tts_engine/llama_local.py:120-157``).  Token strings come from an injected async source
(e.g. the reference's ``remote_backend.generate_tokens_from_api``); the LLM side is out of scope.

Registration (what a maintainer adds)::

    from project_morpheus_b200.adapter import register
    register(registry)            # registry.create("snac_b200", prompt=..., voice=...)
"""
from __future__ import annotations

import asyncio
from dataclasses import dataclass
from typing import Any, AsyncIterator, Callable, Dict, Optional

SAMPLE_RATE = 24000

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
                 token_source: Optional[TokenSource] = None, decoder: Optional[Callable[[AsyncIterator[str]], AsyncIterator[bytes]]] = None,
                 **source_kwargs: Any) -> None:
        self.prompt, self.voice = prompt, voice
        self.use_batching, self.max_batch_chars = use_batching, max_batch_chars
        self._source = token_source
        self._source_kwargs = source_kwargs
        self._decoder = decoder
        self._gen: Optional[AsyncIterator[bytes]] = None
        self._buffer = bytearray()
        self._exhausted = False

    @classmethod
    def configure_token_source(cls, fn: TokenSource) -> None:
        cls._default_source = staticmethod(fn)

    def _ensure_gen(self) -> None:
        if self._gen is None and not self._exhausted:
            src = self._source or type(self)._default_source
            tokens = src(prompt=self.prompt, voice=self.voice, **self._source_kwargs)
            if self._decoder is None:
                from . import speechpipe  # CUDA path; no CPU fallback

                self._decoder = speechpipe.tokens_decoder
            self._gen = self._decoder(tokens).__aiter__()

    async def pull(self, chunk_size: int) -> AudioChunk:
        target = max(0, int(chunk_size))
        self._ensure_gen()
        while len(self._buffer) < target and not self._exhausted:
            assert self._gen is not None
            try:
                self._buffer.extend(await self._gen.__anext__())
            except StopAsyncIteration:
                self._exhausted = True
        if not self._buffer and self._exhausted:
            return AudioChunk(pcm=b"", duration_ms=0.0, eos=True)
        pcm = bytes(self._buffer[:target])
        del self._buffer[:target]
        return AudioChunk(pcm=pcm, duration_ms=len(pcm) / 2 / SAMPLE_RATE * 1000.0,
                          eos=self._exhausted and not self._buffer)

    async def reset(self) -> None:
        gen, self._gen = self._gen, None
        self._buffer.clear()
        self._exhausted = False
        if gen is not None and hasattr(gen, "aclose"):
            try:
                await gen.aclose()
            except Exception:  # noqa: BLE001
                pass


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
