"""N2 (native token ingress, csrc/ingest.cpp) without a GPU: the C++ parser against the golden token-id vectors and
the Python restatement, the native tick scheduler against the Python one and the oracle's per-stream semantics."""
import ctypes as C
import hashlib

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from oracle import speechpipe_ref as sp
from project_morpheus_b200 import _lib, tokens
from project_morpheus_b200.ingest import NativeIngest, NativeTickScheduler
from project_morpheus_b200.scheduler import TickScheduler
from helpers import load_golden


def parse_native(text: str, index: int):
    lib = _lib.load()
    b = text.encode("utf-8", "surrogatepass")
    out = C.c_int64()
    rc = lib.snacb_parse_token(b, len(b), index, C.byref(out))
    assert rc in (0, 1)
    return int(out.value) if rc == 1 else None


def fake_convert(window):
    if len(window) < 7:
        return None
    f = len(window) // 7
    toks = list(window[: 7 * f])
    if any(t < 0 or t > 4096 for t in toks):
        return None
    if f == 1:
        return b""
    h = hashlib.sha256(np.asarray(toks, dtype=np.int32).tobytes()).digest()
    return (h * 128)[:4096]


def fake_batch(windows):
    return [fake_convert(w) for w in windows]


def test_parser_matches_golden_token_ids(ensure_lib):
    for row in load_golden()["g2_token_ids"]:
        assert parse_native(row["text"], row["index"]) == row["id"], row


ASCII_PIECES = ["<custom_token_", ">", " ", "\t", "\n", "-", "+", "_", "0", "1", "7", "42", "4106", "28681", "x", "<", "custom",
                "<custom_token_12>", "\x1c", "\r"]


@settings(max_examples=400, deadline=None)
@given(st.lists(st.sampled_from(ASCII_PIECES), min_size=0, max_size=8), st.integers(min_value=0, max_value=60))
def test_parser_matches_python_restatement_on_ascii(pieces, index):
    text = "".join(pieces)
    tokens.token_id_cache.clear()
    want = tokens.turn_token_into_id(text, index)
    if want is not None and abs(want) > (1 << 61):
        return  # beyond the documented clamp
    assert parse_native(text, index) == want, repr(text)


NUMBER_FIELDS = ["7", " 7 ", "\t7\n", "\x0b7\x0c", "\x1c7", "7\x1f", "\x1d7\x1e", " \x1c 7", "+7", "-7", "4_106", "_7", "7_", "", " ",
                 "0x7", "7.0", "\r28681\r", "1\x1c2"]


@pytest.mark.parametrize("field", NUMBER_FIELDS)
@pytest.mark.parametrize("outer", ["{}", " {} ", "\x1c{}\x1f", "junk{}", "<custom_token_3>{}"])
def test_parser_number_field_whitespace_is_int_whitespace(ensure_lib, field, outer):
    """int() skips ' \\t\\n\\v\\f\\r' only: 0x1c-0x1f inside the number field make the reference return None
    (speechpipe.py:179-181), while str.strip() does remove them around the whole token (speechpipe.py:166)."""
    text = outer.format("<custom_token_" + field + ">")
    for index in (0, 1, 6, 9):
        tokens.token_id_cache.clear()
        assert parse_native(text, index) == tokens.turn_token_into_id(text, index), repr(text)
    if "\x1c7" in field or "7\x1f" in field:
        assert parse_native("<custom_token_" + field + ">", 1) is None


@settings(max_examples=400, deadline=None)
@given(st.lists(st.sampled_from(["0", "1", "7", "42", "4106", " ", "\t", "\x1c", "\x1f", "\r", "_", "+", "-", "x"]), min_size=0, max_size=6),
       st.integers(min_value=0, max_value=60))
def test_parser_matches_python_restatement_on_wrapped_number_fields(pieces, index):
    """Every example is a syntactically complete token: prefix + field + '>' (the open-ended strategy above rarely forms one)."""
    text = "<custom_token_" + "".join(pieces) + ">"
    tokens.token_id_cache.clear()
    want = tokens.turn_token_into_id(text, index)
    if want is not None and abs(want) > (1 << 61):
        return
    assert parse_native(text, index) == want, repr(text)


def dirty_stream(seed, frames):
    rng = np.random.default_rng(seed)
    s = sp.synth_token_strings(seed, frames)
    for pos in sorted(rng.integers(0, len(s), frames).tolist(), reverse=True):
        s.insert(pos, rng.choice(["<custom_token_10>", "junk", "<custom_token_3>", "", " <custom_token_4200> ", "<custom_token_99"]))
    return s


def run_scheduler(sched, streams, seed, evict=None):
    for i in streams:
        sched.add_stream(i)
    cursors = {i: 0 for i in streams}
    got = {i: [] for i in streams}
    rng = np.random.default_rng(seed)
    gone = set()
    while any(cursors[i] < len(streams[i]) for i in streams if i not in gone):
        for i in streams:
            if i in gone:
                continue
            n = int(rng.integers(0, 12))
            chunk = streams[i][cursors[i]: cursors[i] + n]
            cursors[i] += len(chunk)
            sched.push_many(i, chunk)
        sched.tick()
        for i in streams:
            if i not in gone:
                got[i] += sched.pop_audio(i)
        if evict is not None and evict not in gone and cursors[evict] >= 21:
            sched.evict(evict)
            gone.add(evict)
    for i in streams:
        if i not in gone:
            sched.finish(i)
    sched.drain()
    for i in streams:
        if i not in gone:
            got[i] += sched.pop_audio(i)
            assert sched.done(i)
    return got


def test_native_scheduler_matches_python_scheduler_and_oracle(ensure_lib):
    streams = {i: (dirty_stream(i, f) if i % 2 else sp.synth_token_strings(i, f))
               for i, f in enumerate([1, 2, 3, 4, 5, 8, 10, 13, 7, 9, 0, 21])}
    want = {i: list(sp.decode_stream(s, fake_convert)) for i, s in streams.items()}
    py = run_scheduler(TickScheduler(fake_batch), streams, 0)
    nat = run_scheduler(NativeTickScheduler(fake_batch, max_streams=16), streams, 0)
    for i in streams:
        assert nat[i] == py[i] == want[i], i


def test_native_scheduler_golden_chunk_sizes(ensure_lib):
    for row in load_golden()["g3_chunk_sizes"]:
        if not isinstance(row["frames"], int):
            continue
        sched = NativeTickScheduler(fake_batch, max_streams=2)
        got = run_scheduler(sched, {0: sp.synth_token_strings(3, row["frames"])}, 1)[0]
        assert [len(c) for c in got] == row["sizes"], row


def test_native_scheduler_evict_and_slot_reuse(ensure_lib):
    streams = {i: sp.synth_token_strings(40 + i, f) for i, f in enumerate([6, 9, 5, 12])}
    want = {i: list(sp.decode_stream(s, fake_convert)) for i, s in streams.items()}
    sched = NativeTickScheduler(fake_batch, max_streams=4)
    got = run_scheduler(sched, streams, 3, evict=2)
    for i in (0, 1, 3):
        assert got[i] == want[i]
    assert len(got[2]) < len(want[2]) and got[2] == want[2][: len(got[2])]
    sched.add_stream("again")  # the evicted slot starts clean
    sched.push_many("again", streams[0])
    sched.finish("again")
    sched.drain()
    assert sched.pop_audio("again") == want[0]


def test_first_chunk_latch_follows_decode_result(ensure_lib):
    """A rejected first chunk keeps the stream in first-chunk mode: the next accepted token probes again."""
    calls = []

    def decode(windows):
        calls.append([list(w) for w in windows])
        return [None if len(calls) == 1 else fake_convert(w) for w in windows]

    sched = NativeTickScheduler(decode, max_streams=1)
    sched.add_stream(0)
    s = sp.synth_token_strings(5, 2)
    sched.push_many(0, s[:9])
    assert sched.tick() == 1 and len(calls[0][0]) == 7
    assert sched.tick() == 1 and len(calls[1][0]) == 7 and calls[1][0] != calls[0][0]
    assert sched.tick() == 0  # latched; the ninth token does not complete a frame


def test_ingest_argument_checks(ensure_lib):
    ing = NativeIngest(2)
    with pytest.raises(_lib.SnacbError):
        ing.push([5], ["<custom_token_11>"])
    ing.finish(1)
    with pytest.raises(_lib.SnacbError):
        ing.push([1], ["<custom_token_11>"])
    tok, ntok, owner = ing.tick()
    assert len(ntok) == 0 and ing.done(1) and not ing.done(0)
    assert ing.stats() == {"accepted": 0, "rejected": 0, "windows": 0}
    ing.close()


class FakeEngine:
    """submit_windows / wait_windows with the deterministic stand-in decode (two tickets, like the C ABI)."""

    def __init__(self):
        self.slots, self.next = {}, 0
        self.max_inflight = 0

    def submit_windows(self, tokens, ntok=None, noise="philox", seed=0, keys=None):
        tok = np.array(tokens, dtype=np.int32, copy=True)
        lens = [tok.shape[1]] * tok.shape[0] if ntok is None else [int(x) for x in ntok]
        t = self.next
        assert t not in self.slots, "two ticks already in flight"
        self.slots[t] = (tok, lens)
        self.next ^= 1
        self.max_inflight = max(self.max_inflight, len(self.slots))
        return t

    def wait_windows(self, ticket):
        tok, lens = self.slots.pop(ticket)
        pcm = np.zeros((len(lens), 2048), dtype=np.int16)
        st_ = np.zeros((len(lens),), dtype=np.int32)
        for i, n in enumerate(lens):
            r = fake_convert(tok[i, :n].tolist())
            if r is None:
                st_[i] = _lib.WIN_REJECTED
            elif len(r) == 0:
                st_[i] = _lib.WIN_EMPTY
            else:
                pcm[i] = np.frombuffer(r, dtype="<i2")
        return pcm, st_


def test_pipelined_native_scheduler_keeps_per_stream_results(ensure_lib):
    """Pipelined ticks (submit now, deliver the previous tick): same chunks per stream as the oracle's per-stream
    semantics, incl. dirty streams, a rejected first chunk, an evicted stream and slot reuse."""
    streams = {i: (dirty_stream(i, f) if i % 2 else sp.synth_token_strings(i, f))
               for i, f in enumerate([1, 2, 3, 4, 5, 8, 10, 13, 7, 9, 0, 21])}
    streams[12] = ["<custom_token_9000>"] * 3 + sp.synth_token_strings(77, 6)  # out-of-range ids: first chunks are rejected
    want = {i: list(sp.decode_stream(s, fake_convert)) for i, s in streams.items()}
    eng = FakeEngine()
    sched = NativeTickScheduler(max_streams=16, engine=eng)
    got = run_scheduler(sched, streams, 0)
    assert eng.max_inflight >= 1
    for i in streams:
        assert got[i] == want[i], i
    # evict with a window in flight: nothing of the evicted stream is delivered afterwards, its slot starts clean
    sched2 = NativeTickScheduler(max_streams=2, engine=FakeEngine())
    sched2.add_stream("a")
    sched2.push_many("a", streams[5])
    sched2.tick()
    sched2.evict("a")
    sched2.add_stream("b")
    sched2.push_many("b", streams[3])
    sched2.finish("b")
    sched2.drain()
    assert sched2.pop_audio("b") == want[3]
