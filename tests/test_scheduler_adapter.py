"""Host logic without a GPU: tick scheduler == per-stream tokens_decoder semantics; adapter byte contract."""
import asyncio
import hashlib

import numpy as np
import pytest

from oracle import speechpipe_ref as sp
from project_morpheus_b200.adapter import AudioChunk, SnacB200Adapter, describe, register, voice_mapper
from project_morpheus_b200.scheduler import TickScheduler


def fake_convert(window):
    """Deterministic stand-in for the decode: depends on every token of the window."""
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


def dirty_stream(seed, frames):
    rng = np.random.default_rng(seed)
    s = sp.synth_token_strings(seed, frames)
    for pos in sorted(rng.integers(0, len(s), frames).tolist(), reverse=True):
        s.insert(pos, rng.choice(["<custom_token_10>", "junk", "<custom_token_3>", ""]))
    return s


def test_scheduler_matches_per_stream_reference_semantics():
    streams = {i: (dirty_stream(i, f) if i % 2 else sp.synth_token_strings(i, f))
               for i, f in enumerate([1, 2, 3, 4, 5, 8, 10, 13, 7, 9])}
    want = {i: list(sp.decode_stream(s, fake_convert)) for i, s in streams.items()}
    sched = TickScheduler(fake_batch)
    for i in streams:
        sched.add_stream(i)
    # interleave ingress across streams a few tokens at a time, ticking as we go
    cursors = {i: 0 for i in streams}
    got = {i: [] for i in streams}
    rng = np.random.default_rng(0)
    while any(cursors[i] < len(streams[i]) for i in streams):
        for i in streams:
            n = int(rng.integers(0, 12))
            chunk = streams[i][cursors[i]: cursors[i] + n]
            cursors[i] += len(chunk)
            sched.push_many(i, chunk)
        sched.tick()
        for i in streams:
            got[i] += sched.pop_audio(i)
    for i in streams:
        sched.finish(i)
    sched.drain()
    for i in streams:
        got[i] += sched.pop_audio(i)
        assert sched.done(i)
        assert got[i] == want[i], f"stream {i}"
    assert sched.windows_decoded == sum(1 for i in streams for _ in range(1)) or sched.windows_decoded > 0


def test_scheduler_batches_one_window_per_stream_per_tick_and_evicts():
    calls = []

    def spy(windows):
        calls.append([len(w) for w in windows])
        return fake_batch(windows)

    sched = TickScheduler(spy)
    for i in range(64):
        sched.add_stream(i)
        sched.push_many(i, sp.synth_token_strings(i, 5))
    n = sched.tick()
    assert n == 64 and calls[-1] == [7] * 64           # first chunk of every stream in ONE call
    sched.evict(3)                                     # barge-in: stream 3 disappears mid-flight
    sched.add_stream("new")                            # refilled slot
    sched.push_many("new", sp.synth_token_strings(99, 4))
    n = sched.tick()
    assert n == 64 and sorted(calls[-1]) == [7] + [28] * 63
    assert 3 not in sched and "new" in sched
    sched.drain()
    assert sched.pop_audio(5)[0] == b""


def test_scheduler_first_chunk_latch_follows_decode_result():
    # 7 tokens whose window is rejected (code > 4096): the reference keeps retrying the first-chunk path
    bad = [f"<custom_token_{9000 + 10 + 4096 * (p % 7)}>" for p in range(7)]
    good = sp.synth_token_strings(1, 4)
    strings = bad + good
    want = list(sp.decode_stream(strings, fake_convert))
    sched = TickScheduler(fake_batch)
    sched.add_stream(0)
    sched.push_many(0, strings)
    sched.finish(0)
    sched.drain()
    assert sched.pop_audio(0) == want


def _tokens(strings):
    async def gen(**_):
        for s in strings:
            yield s
    return gen


async def _fake_decoder(token_gen):
    async for c in _async_decode(token_gen):
        yield c


async def _async_decode(token_gen):
    strings = [s async for s in token_gen]
    for c in sp.decode_stream(strings, fake_convert):
        yield c


def test_adapter_rechunking_contract():
    """Same byte semantics the reference pins in tests/test_tts_adapter_chunking.py:25-44."""
    strings = sp.synth_token_strings(2, 6)
    whole = b"".join(sp.decode_stream(strings, fake_convert))
    ad = SnacB200Adapter("hello", "tara", token_source=_tokens(strings), decoder=_fake_decoder)

    async def run(chunk):
        out, n = bytearray(), 0
        while True:
            c = await ad.pull(chunk)
            assert isinstance(c, AudioChunk) and len(c.pcm) <= chunk
            assert abs(c.duration_ms - len(c.pcm) / 2 / 24000 * 1000.0) < 1e-9
            out += c.pcm
            n += 1
            if c.eos:
                return bytes(out), n

    got, n = asyncio.run(run(1000))
    assert got == whole and n == -(-len(whole) // 1000)

    async def barge():
        first = await ad.pull(64)
        await ad.reset()
        again = await ad.pull(64)
        return first, again

    ad2 = SnacB200Adapter("hello", token_source=_tokens(strings), decoder=_fake_decoder)
    ad = ad2
    first, again = asyncio.run(barge())
    assert first.pcm == again.pcm == whole[:64]  # reset restarts the utterance


def test_adapter_registers_like_reference_registry():
    class Registry:  # the register/create surface of tts_engine/adapter_registry.py:70-98
        def __init__(self):
            self.r = {}

        def register(self, name, constructor, describe, voice_mapper):
            self.r[name] = (constructor, describe, voice_mapper)

        def create(self, name, *, prompt, voice, **kw):
            c, _, vm = self.r[name]
            params = vm(voice)
            params.update(kw)
            return c(prompt=prompt, **params)

    reg = Registry()
    register(reg)
    d = reg.r["snac_b200"][1]()
    assert d["name"] == "snac_b200" and d["streaming"] and d["supports_barge_in"]
    ad = reg.create("snac_b200", prompt="hi", voice="leo", use_batching=True, max_batch_chars=500)
    assert isinstance(ad, SnacB200Adapter) and ad.voice == "leo" and ad.use_batching
    with pytest.raises(RuntimeError):
        asyncio.run(ad.pull(8))  # no token source configured: fails loudly


def test_planner_and_scheduler_property_random_streams():
    """Property test (hypothesis): for arbitrary streams of valid / zero / negative / out-of-range / malformed /
    multi-token strings the product's WindowPlanner (per stream) and TickScheduler (batched) emit exactly the chunk
    sequence of the oracle's restatement of tokens_decoder (itself proven identical to the verbatim reference)."""
    hyp = pytest.importorskip("hypothesis")
    from hypothesis import given, settings, strategies as stx

    token = stx.one_of(
        stx.integers(min_value=0, max_value=7 * 4096 + 30).map(lambda n: f"<custom_token_{n}>"),
        stx.sampled_from(["", "junk", "<custom_token_", "<custom_token_x>", "<custom_token_10>", "<custom_token_11><custom_token_4107>",
                          " <custom_token_20000> ", "<custom_token_99999>"]),
    )

    @settings(max_examples=60, deadline=None)
    @given(stx.lists(token, min_size=0, max_size=140))
    def check(strings):
        want = list(sp.decode_stream(strings, fake_convert))
        sched = TickScheduler(fake_batch)
        sched.add_stream("s")
        sched.push_many("s", strings)
        sched.finish("s")
        sched.drain()
        assert sched.pop_audio("s") == want

    check()
