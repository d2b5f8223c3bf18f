"""Oracle: independent restatement of the reference's ``speechpipe`` semantics.

TEST INFRASTRUCTURE ONLY - see ``oracle/__init__.py``.

Follows ``/root/reference/Morpheus_Client/tts_engine/speechpipe.py``:
  * ``parse_custom_token``  <- ``turn_token_into_id``      ``speechpipe.py:146-189``
  * ``split_levels``        <- de-interleave + validator    ``speechpipe.py:72-111``
  * ``window_to_pcm``       <- ``convert_to_audio``         ``speechpipe.py:64-137``
  * ``StreamWindower``      <- ``tokens_decoder``           ``speechpipe.py:191-293``
  * ``drop_empty_in_fives`` <- ``tokens_decoder_sync``      ``speechpipe.py:295-337``

It is written from the behaviour table in SURVEY.md Appendix B (Q1-Q12), not
from the reference text, and ``tests/test_oracle_vs_reference.py`` proves it
byte-identical to the verbatim reference file on seeded and adversarial
streams whenever ``/root/reference`` is mounted.  It exists because the GPU box
has no ``/root/reference``.
"""
from __future__ import annotations

from typing import Callable, Iterable, Iterator, List, Optional, Sequence, Tuple

import numpy as np

PREFIX = "<custom_token_"
TOKENS_PER_FRAME = 7
# position inside a 7-token frame -> (level, index inside the frame's share of that level)
SLOT_TO_LEVEL = ((0, 0), (1, 0), (2, 0), (2, 1), (1, 1), (2, 2), (2, 3))
CODES_PER_FRAME = (1, 2, 4)
SLICE_LO, SLICE_HI = 2048, 4096


def parse_custom_token(text: str, index: int) -> Optional[int]:
    """Q3: the LAST ``<custom_token_N>`` of ``text`` -> ``N - 10 - 4096*(index%7)``; no range check."""
    if PREFIX not in text:
        return None
    tail = text.strip()
    at = tail.rfind(PREFIX)
    if at < 0:
        return None
    tail = tail[at:]
    if not tail.endswith(">"):
        return None
    try:
        n = int(tail[len(PREFIX):-1])
    except ValueError:
        return None
    return n - 10 - 4096 * (index % TOKENS_PER_FRAME)


def split_levels(tokens: Sequence[int]) -> Optional[Tuple[np.ndarray, np.ndarray, np.ndarray, bool]]:
    """De-interleave whole frames into the three RVQ levels (strides 4/2/1).

    Returns ``None`` for fewer than 7 tokens, else ``(c0, c1, c2, ok)`` where ``ok`` is the
    reference validator's verdict: every code in ``[0, 4096]`` (Q1: 4096 passes).
    """
    if len(tokens) < TOKENS_PER_FRAME:
        return None
    frames = len(tokens) // TOKENS_PER_FRAME
    t = np.asarray(tokens[: frames * TOKENS_PER_FRAME], dtype=np.int64)
    t = t.astype(np.int32).reshape(frames, TOKENS_PER_FRAME)  # reference builds an int32 tensor
    levels = [np.zeros(frames * n, dtype=np.int32) for n in CODES_PER_FRAME]
    for slot, (lvl, k) in enumerate(SLOT_TO_LEVEL):
        levels[lvl][k :: CODES_PER_FRAME[lvl]] = t[:, slot]
    ok = all(bool(((c >= 0) & (c <= 4096)).all()) for c in levels)
    return levels[0], levels[1], levels[2], ok


def pcm16_bytes(audio: np.ndarray) -> bytes:
    """Q7: ``trunc(x * 32767)`` in float32, little-endian int16, no clip, no rounding."""
    return (np.asarray(audio, dtype=np.float32) * np.float32(32767)).astype(np.int16).tobytes()


DecodeFn = Callable[[np.ndarray, np.ndarray, np.ndarray], np.ndarray]  # codes -> float32[2048*F]


def window_to_pcm(tokens: Sequence[int], decode: DecodeFn) -> Optional[bytes]:
    """One window -> ``None`` (skipped) | ``b''`` (F == 1) | 4096 bytes."""
    lv = split_levels(tokens)
    if lv is None or not lv[3]:
        return None
    audio = decode(lv[0], lv[1], lv[2])
    return pcm16_bytes(audio[SLICE_LO:SLICE_HI])


class StreamWindower:
    """Per-stream sliding window state machine (``tokens_decoder``), decode-agnostic.

    ``push(text)`` returns the token window to decode now (or ``None``); the caller
    reports back through ``decoded(result_is_none)`` because the first-chunk latch only
    flips when the first decode is not skipped.  ``finish()`` returns the end-of-stream window.
    """

    FIRST, SHORT, LONG = 7, 28, 49

    def __init__(self) -> None:
        self.ids: List[int] = []
        self.accepted = 0
        self.first_done = False
        self._awaiting_first = False

    def push(self, text: str) -> Optional[List[int]]:
        tid = parse_custom_token(text, self.accepted)
        if tid is None or tid <= 0:  # Q2: dropped, slot index does not advance
            return None
        self.ids.append(tid)
        self.accepted += 1
        if not self.first_done:
            if self.accepted >= self.FIRST:
                self._awaiting_first = True
                return self.ids[-self.FIRST:]
            return None
        if self.accepted % TOKENS_PER_FRAME:
            return None
        if len(self.ids) >= self.LONG:
            return self.ids[-self.LONG:]
        if len(self.ids) >= self.SHORT:
            return self.ids[-self.SHORT:]
        return None

    def decoded(self, result_is_none: bool) -> None:
        if self._awaiting_first:
            self._awaiting_first = False
            if not result_is_none:
                self.first_done = True

    def finish(self) -> Optional[List[int]]:
        n = len(self.ids)
        if n >= self.LONG:
            return self.ids[-self.LONG:]
        if n >= self.SHORT:
            return self.ids[-self.SHORT:]
        if n >= TOKENS_PER_FRAME:  # Q6: pad with copies of the last token up to 28
            return self.ids + [self.ids[-1]] * (self.SHORT - n)
        return None


def decode_stream(texts: Iterable[str], convert: Callable[[List[int]], Optional[bytes]]) -> Iterator[bytes]:
    """``tokens_decoder`` as a plain generator over an iterable of token strings."""
    sw = StreamWindower()
    for text in texts:
        win = sw.push(text)
        if win is None:
            continue
        out = convert(win)
        sw.decoded(out is None)
        if out is not None:
            yield out
    win = sw.finish()
    if win is not None:
        out = convert(win)
        if out is not None:
            yield out


def drop_empty_in_fives(chunks: Iterable[bytes]) -> Iterator[bytes]:
    """Q10: ``tokens_decoder_sync`` drops falsy chunks and releases audio in groups of five."""
    held: List[bytes] = []
    for c in chunks:
        if not c:
            continue
        held.append(c)
        if len(held) >= 5:
            yield from held
            held = []
    yield from held


# --------------------------------------------------------------------------- synthetic streams
def synth_codes(stream: int, frames: int) -> np.ndarray:
    """SURVEY 8(d) token recipe: PCG64(1234+stream), codes U{1..4095}, shape [frames*7]."""
    rng = np.random.Generator(np.random.PCG64(1234 + stream))
    return rng.integers(1, 4096, size=frames * TOKENS_PER_FRAME, dtype=np.int64)


def synth_token_strings(stream: int, frames: int) -> List[str]:
    codes = synth_codes(stream, frames)
    return [f"<custom_token_{int(c) + 10 + 4096 * (p % TOKENS_PER_FRAME)}>" for p, c in enumerate(codes)]
