"""Host-side token logic of the ``speechpipe`` interface (no GPU, no engine import).

* ``turn_token_into_id``  <- reference ``Morpheus_Client/tts_engine/speechpipe.py:146-189``
* ``WindowPlanner``       <- the control flow of reference ``tokens_decoder`` (``:191-293``), separated
  from the decode so a tick scheduler can batch the windows of many streams into one launch.
"""
from __future__ import annotations

from typing import List, Optional

TOKENS_PER_FRAME = 7

CUSTOM_TOKEN_PREFIX = "<custom_token_"
token_id_cache = {}
MAX_CACHE_SIZE = 10000


def turn_token_into_id(token_string: str, index: int) -> Optional[int]:
    """Last ``<custom_token_N>`` in the string -> ``N - 10 - (index % 7) * 4096`` (no range check)."""
    slot = index % TOKENS_PER_FRAME
    key = (token_string, slot)
    hit = token_id_cache.get(key, key)
    if hit is not key:
        return hit
    if CUSTOM_TOKEN_PREFIX not in token_string:
        return None
    text = token_string.strip()
    start = text.rfind(CUSTOM_TOKEN_PREFIX)
    if start < 0 or not text.endswith(">"):
        return None
    try:
        value = int(text[start + len(CUSTOM_TOKEN_PREFIX):-1]) - 10 - slot * 4096
    except ValueError:
        return None
    if len(token_id_cache) < MAX_CACHE_SIZE:
        token_id_cache[key] = value
    return value


# ----------------------------------------------------------------------------- per-stream driver
FIRST_WINDOW = 7    # first audio after one frame (its slice is empty: b'')
SHORT_WINDOW = 28   # 4 frames
LONG_WINDOW = 49    # 7 frames


class WindowPlanner:
    """Sliding-window state of one stream (the control flow of the reference ``tokens_decoder``),
    separated from the decode so a tick scheduler can batch windows of many streams."""

    __slots__ = ("buffer", "count", "first_done", "_first_pending")

    def __init__(self) -> None:
        self.buffer: List[int] = []
        self.count = 0
        self.first_done = False
        self._first_pending = False

    def push(self, token_string: str) -> Optional[List[int]]:
        """Feed one token string; returns the window to decode now, if any."""
        token = turn_token_into_id(token_string, self.count)
        if token is None or token <= 0:
            return None
        self.buffer.append(token)
        self.count += 1
        if not self.first_done:
            if self.count >= FIRST_WINDOW:
                self._first_pending = True
                return self.buffer[-FIRST_WINDOW:]
            return None
        if self.count % TOKENS_PER_FRAME != 0:
            return None
        return self._steady_window()

    def _steady_window(self) -> Optional[List[int]]:
        if len(self.buffer) >= LONG_WINDOW:
            return self.buffer[-LONG_WINDOW:]
        if len(self.buffer) >= SHORT_WINDOW:
            return self.buffer[-SHORT_WINDOW:]
        return None

    def result(self, audio: Optional[bytes]) -> None:
        """Report the outcome of the window ``push`` returned (latches the first chunk)."""
        if self._first_pending:
            self._first_pending = False
            if audio is not None:
                self.first_done = True

    def flush(self) -> Optional[List[int]]:
        """End of stream: last 49 / last 28 / (>= 7 tokens) padded to 28 with the last token."""
        win = self._steady_window()
        if win is not None:
            return win
        if len(self.buffer) >= TOKENS_PER_FRAME:
            return self.buffer + [self.buffer[-1]] * (SHORT_WINDOW - len(self.buffer))
        return None
