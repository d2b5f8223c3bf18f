"""Oracle for row N3 (PCM egress): numpy restatement of the reference stitcher and RIFF header.

TEST INFRASTRUCTURE ONLY - see ``oracle/__init__.py``.
* ``stitch``       <- ``Morpheus_Client/orchestrator/stitcher.py:10-79`` as a synchronous generator over
  ``(pcm_bytes, eos)`` pairs yielding ``(pcm_bytes, eos)``
* ``riff_header``  <- ``Morpheus_Client/server.py:50-70``
Pinned against the verbatim reference by ``tests/golden/make_golden_egress.py`` (committed vectors) and, where the
reference tree is mounted, by ``tests/test_egress.py::test_oracle_matches_verbatim_reference``.
"""
from __future__ import annotations

import struct
from typing import Iterable, Iterator, Tuple

import numpy as np


def riff_header(sample_rate: int = 24000) -> bytes:
    return struct.pack("<4sI4s4sIHHIIHH4sI", b"RIFF", 0xFFFFFFFF, b"WAVE", b"fmt ", 16, 1, 1, sample_rate, sample_rate * 2, 2, 16,
                       b"data", 0xFFFFFFFF)


def stitch(chunks: Iterable[Tuple[bytes, bool]], sample_rate: int, overlap_ms: float = 0.0) -> Iterator[Tuple[bytes, bool]]:
    tail = np.zeros(0, dtype=np.int16)
    overlap = int(overlap_ms * sample_rate / 1000.0)
    for data, eos in chunks:
        pcm = np.frombuffer(data, dtype=np.int16)
        if tail.size:
            ov = min(overlap, tail.size, pcm.size) if overlap > 0 else 0
            if ov:
                fade_out = tail[-ov:] * np.linspace(1.0, 0.0, ov, endpoint=False)
                fade_in = pcm[:ov] * np.linspace(0.0, 1.0, ov, endpoint=False)
                pcm = np.concatenate([tail[:-ov], fade_out + fade_in, pcm[ov:]])
            else:
                pcm = np.concatenate([tail, pcm])
        if eos:
            yield pcm.astype("<i2").tobytes(), True
            tail = np.zeros(0, dtype=np.int16)
            return
        if overlap > 0:
            if pcm.size <= overlap:
                tail = pcm
                continue
            out, tail = pcm[:-overlap], pcm[-overlap:]
        else:
            out, tail = pcm, np.zeros(0, dtype=np.int16)
        yield out.astype("<i2").tobytes(), False
    if tail.size:
        yield tail.astype("<i2").tobytes(), True
