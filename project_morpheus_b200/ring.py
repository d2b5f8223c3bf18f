"""Per-stream PCM ring between the decode ticks and ``TTSAdapter.pull`` (adapter side).

The reference re-chunks by extending one ``bytearray`` and deleting its head on every ``pull``
(``/root/reference/Morpheus_Client/tts_engine/llama_local.py:131-150``); with the default chunk ladder of 8..64 bytes
(``orchestrator/chunk_ladder.py:7``) that is a memmove of the whole backlog per pull.  Here decoded chunks are kept as
the immutable ``bytes`` objects the decode produced, in a deque; ``read(n)`` hands out views of the head segment(s) and
only ever copies the ``n`` bytes it returns.  Same byte stream, O(n) per pull regardless of backlog.
"""
from __future__ import annotations

from collections import deque
from typing import Deque


class PcmRing:
    __slots__ = ("_segs", "_off", "_size", "written", "read_bytes")

    def __init__(self) -> None:
        self._segs: Deque[bytes] = deque()
        self._off = 0      # read offset into the head segment
        self._size = 0     # unread bytes
        self.written = 0
        self.read_bytes = 0

    def __len__(self) -> int:
        return self._size

    def write(self, data: bytes) -> None:
        if data:
            self._segs.append(data)
            self._size += len(data)
            self.written += len(data)

    def read(self, n: int) -> bytes:
        """Up to ``n`` bytes from the head (fewer only when the ring runs dry)."""
        n = min(max(0, int(n)), self._size)
        if n == 0:
            return b""
        head = self._segs[0]
        if len(head) - self._off >= n:  # the common case: one segment serves the pull
            out = head[self._off:self._off + n]
            self._off += n
            if self._off == len(head):
                self._segs.popleft()
                self._off = 0
        else:
            parts, need = [], n
            while need:
                head = self._segs[0]
                take = min(need, len(head) - self._off)
                parts.append(memoryview(head)[self._off:self._off + take])
                self._off += take
                need -= take
                if self._off == len(head):
                    self._segs.popleft()
                    self._off = 0
            out = b"".join(parts)
        self._size -= n
        self.read_bytes += n
        return out

    def clear(self) -> None:
        self._segs.clear()
        self._off = 0
        self._size = 0
