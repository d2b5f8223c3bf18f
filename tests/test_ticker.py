"""The tick batcher behind the reference's seam (no GPU): many adapters pulled concurrently under one event loop through
the reference's Orchestrator pattern coalesce into few batched decode calls, with per-stream bytes identical to the
serial per-stream path; per-stream noise keys; error isolation; the PCM ring; the product's tokens_decoder_sync."""
import asyncio
import hashlib
import importlib
import os
import sys

import numpy as np
import pytest

from oracle import speechpipe_ref as sp
from project_morpheus_b200 import ticker as tk
from project_morpheus_b200.adapter import SnacB200Adapter
from project_morpheus_b200.ring import PcmRing

REF_ROOT = "/root/reference"


def fake_convert(window):
    if len(window) < 7:
        return None
    f = len(window) // 7
    toks = list(window[: 7 * f])
    if any(t < 0 or t > 4096 for t in toks):
        return None
    if f == 1:
        return b""
    h = hashlib.sha256(np.asarray(toks, dtype=np.int64).tobytes()).digest()
    return (h * 128)[:4096]


class FakeGpuRing:
    """Host stand-in with the interface of egress.GpuPcmRing (the real one needs a CUDA device; GPU tests cover it)."""

    overlap_samples = 0
    ring_samples = 1 << 15

    def __init__(self, n):
        self.bufs = [bytearray() for _ in range(n)]
        self._free = list(range(n - 1, -1, -1))

    def acquire(self):
        if not self._free:
            from project_morpheus_b200 import _lib
            raise _lib.SnacbError("all slots are in use")
        return self._free.pop()

    def release(self, slot):
        self.reset(slot)
        self._free.append(slot)

    def available(self, slot):
        return len(self.bufs[slot])

    def room(self, slot):
        return self.ring_samples - len(self.bufs[slot]) // 2

    def read(self, slot, nbytes):
        n = min(nbytes // 2 * 2, len(self.bufs[slot]) // 2 * 2)
        out = bytes(self.bufs[slot][:n])
        del self.bufs[slot][:n]
        return out

    def reset(self, slot):
        self.bufs[slot].clear()

    def flush(self, slot):
        pass


@pytest.fixture()
def speechpipe(monkeypatch):
    """The product module on a CPU host, its engine call replaced by a deterministic stand-in that records every call."""
    monkeypatch.setenv("SNACB_RANDOM_INIT", "0")
    monkeypatch.setenv("SNACB_NOISE", "philox")
    sys.modules.pop("project_morpheus_b200.speechpipe", None)
    mod = importlib.import_module("project_morpheus_b200.speechpipe")
    calls = []

    def fake_batch(windows, noise=None, keys=None, errors="none"):
        calls.append((len(windows), None if keys is None else [int(k) for k in keys]))
        out = []
        for w in windows:
            if len(w) >= 7 and any(t == 4096 for t in w[: 7 * (len(w) // 7)]):
                out.append(IndexError("index out of range in self") if errors == "values" else None)
            else:
                out.append(fake_convert(w))
        return out

    ring = FakeGpuRing(128)

    def fake_ring_batch(windows, slots, keys=None, errors="none", ring_=None):
        calls.append((len(windows), None if keys is None else [int(k) for k in keys]))
        out = []
        for w, sl in zip(windows, slots):
            r = fake_batch([w], keys=None if keys is None else keys[:1], errors=errors)[0]
            calls.pop()
            if isinstance(r, bytes):
                ring.bufs[sl] += r
                r = len(r)
            out.append(r)
        return out

    monkeypatch.setattr(mod, "convert_to_audio_batch", fake_batch)
    monkeypatch.setattr(mod, "convert_to_ring_batch", fake_ring_batch)
    monkeypatch.setattr(mod, "get_ring", lambda: ring)
    monkeypatch.setattr(mod, "_ticker", None)
    mod.calls = calls
    mod.fake_ring = ring
    yield mod
    sys.modules.pop("project_morpheus_b200.speechpipe", None)


def _source(strings, gap=0):
    async def gen(**_):
        for i, s in enumerate(strings):
            if gap and i % gap == 0:
                await asyncio.sleep(0)  # tokens trickle in like an SSE stream
            yield s
    return gen


async def _orchestrate(adapter, ladder_start=0):
    """Drive one adapter the way the reference does.  With /root/reference present this IS the reference's
    Orchestrator.stream (orchestrator/core.py:74-125) with its default chunk ladder; elsewhere the same pull loop."""
    if os.path.isdir(REF_ROOT):
        if REF_ROOT not in sys.path:
            sys.path.insert(0, REF_ROOT)
        from Morpheus_Client.orchestrator.buffer import PlaybackBuffer
        from Morpheus_Client.orchestrator.core import Orchestrator

        orch = Orchestrator(adapter, PlaybackBuffer(capacity_ms=1000.0))
        out = bytearray()
        async for chunk in orch.stream():
            out += chunk.pcm
        return bytes(out)
    out, sizes, k = bytearray(), [8, 12, 16, 24, 32, 48, 64], ladder_start
    while True:
        chunk = await adapter.pull(sizes[k % len(sizes)])
        out += chunk.pcm
        k += 1
        if chunk.eos:
            return bytes(out)


def test_concurrent_adapters_coalesce_into_ticks(speechpipe):
    """96 requests, each behind its own adapter and orchestrator (core.py:89-117, registry call adapter_registry.py:90-98):
    ticks << windows, every tick is ONE batched engine call, bytes equal the serial per-stream decode."""
    n = 96
    frames = [4 + (i % 9) for i in range(n)]
    streams = [sp.synth_token_strings(i, f) for i, f in enumerate(frames)]
    want = [b"".join(sp.decode_stream(s, fake_convert)) for s in streams]
    n_windows = sum(len(list(sp.decode_stream(s, lambda w: b"x" if len(w) > 7 else b""))) for s in streams)

    async def main():
        ads = [SnacB200Adapter("p", "tara", token_source=_source(s, gap=7), seed=1000 + i) for i, s in enumerate(streams)]
        return await asyncio.gather(*[_orchestrate(a, i) for i, a in enumerate(ads)])

    got = asyncio.run(main())
    assert got == want
    t = speechpipe.get_ticker()
    st = t.stats()
    assert st["windows"] == n_windows and len(speechpipe.calls) == st["ticks"]      # one engine call per tick
    assert st["ticks"] <= 2 * max(frames) and st["ticks"] * 8 < st["windows"], st    # not one call per window
    assert st["max_tick"] >= n // 2, st
    # serial path (ticker off): same bytes, one call per window
    speechpipe.calls.clear()

    async def serial(i):
        return b"".join([c async for c in speechpipe.tokens_decoder(_source(streams[i])(), ticker=False, stream_key=1000 + i)])

    assert [asyncio.run(serial(i)) for i in (0, 17, 95)] == [want[0], want[17], want[95]]
    assert all(c[0] == 1 for c in speechpipe.calls)


def test_noise_keys_depend_on_the_stream_only(speechpipe):
    """Window w of a stream with seed s is keyed window_key(s, w) whatever else runs beside it."""
    strings = sp.synth_token_strings(5, 8)

    async def one(seed, crowd):
        speechpipe.calls.clear()
        jobs = [_drain(SnacB200Adapter("p", token_source=_source(strings, gap=3), seed=seed))]
        jobs += [_drain(SnacB200Adapter("p", token_source=_source(sp.synth_token_strings(50 + j, 6), gap=2))) for j in range(crowd)]
        await asyncio.gather(*jobs)
        return sorted(k for _, keys in speechpipe.calls for k in keys)

    async def _drain(ad):
        while not (await ad.pull(4096)).eos:
            pass

    alone = asyncio.run(one(77, 0))
    n_win = len(alone)
    assert alone == sorted(tk.window_key(77, w) for w in range(n_win))
    crowded = asyncio.run(one(77, 20))
    assert set(alone) <= set(crowded)
    assert tk.window_key(77, 0) != tk.window_key(78, 0) and tk.window_key(77, 0) != tk.window_key(77, 1)


def test_error_in_one_stream_does_not_touch_the_others(speechpipe):
    good = sp.synth_token_strings(3, 6)
    bad = list(good)
    bad[8] = f"<custom_token_{4096 + 10 + 4096 * (8 % 7)}>"  # code 4096: passes the validator, embedding raises (Q1)

    async def main():
        a, b = SnacB200Adapter("p", token_source=_source(good, gap=2)), SnacB200Adapter("p", token_source=_source(bad, gap=2))

        async def drain(ad):
            out = bytearray()
            while True:
                c = await ad.pull(1024)
                out += c.pcm
                if c.eos:
                    return bytes(out)

        ra, rb = await asyncio.gather(drain(a), drain(b), return_exceptions=True)
        return ra, rb

    ra, rb = asyncio.run(main())
    assert ra == b"".join(sp.decode_stream(good, fake_convert))
    assert isinstance(rb, IndexError)  # the reference's tokens_decoder dies with the same IndexError


def test_barge_in_reset_restarts_and_cancels_the_pump(speechpipe):
    strings = sp.synth_token_strings(9, 7)
    whole = b"".join(sp.decode_stream(strings, fake_convert))

    async def main():
        ad = SnacB200Adapter("p", token_source=_source(strings, gap=1), seed=4)
        first = await ad.pull(100)
        pump = ad._pump
        await ad.reset()
        assert pump.done() and ad._pump is None and len(ad._ring) == 0
        out = bytearray()
        while True:
            c = await ad.pull(333)
            assert len(c.pcm) <= 333
            out += c.pcm
            if c.eos:
                return first.pcm, bytes(out)

    first, again = asyncio.run(main())
    assert first == whole[:100] and again == whole


def test_product_tokens_decoder_sync_matches_reference_grouping(speechpipe):
    """a6: the PRODUCT's tokens_decoder_sync (speechpipe.py:295-337 semantics: empty chunks dropped, groups of five)."""
    for stream, frames in ((77, 9), (78, 1), (79, 3), (80, 14), (81, 23)):
        strings = sp.synth_token_strings(stream, frames)
        if stream == 80:  # dirty stream: junk, zero ids, a rejected window
            strings = strings[:20] + ["junk", "<custom_token_10>", ""] + strings[20:]

        async def go():
            return [c async for c in speechpipe.tokens_decoder_sync(_source(strings, gap=5)())]

        got = asyncio.run(go())
        want = list(sp.drop_empty_in_fives(sp.decode_stream(strings, fake_convert)))
        assert got == want and all(got), (stream, frames)


@pytest.mark.needs_reference
@pytest.mark.skipif(not os.path.isdir(REF_ROOT), reason="reference tree not mounted")
def test_product_tokens_decoder_sync_matches_verbatim_reference(speechpipe):
    """Same strings through the VERBATIM reference module (its decode replaced by the same stand-in)."""
    from oracle import ref_loader

    ref = ref_loader.load_reference_speechpipe()
    ref.convert_to_audio = lambda multiframe, count: fake_convert(list(multiframe))
    strings = sp.synth_token_strings(12, 11)
    import contextlib, io

    async def run(mod):
        return [c async for c in mod.tokens_decoder_sync(_source(strings)())]

    with contextlib.redirect_stdout(io.StringIO()):
        want = asyncio.run(run(ref))
    assert asyncio.run(run(speechpipe)) == want


def test_pcm_ring_is_a_byte_fifo():
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=200, deadline=None)
    @given(st.lists(st.one_of(st.binary(min_size=0, max_size=40).map(lambda b: ("w", b)),
                              st.integers(min_value=0, max_value=64).map(lambda n: ("r", n))), max_size=60))
    def check(ops):
        ring, model = PcmRing(), bytearray()
        for op, arg in ops:
            if op == "w":
                ring.write(arg)
                model += arg
            else:
                got = ring.read(arg)
                assert got == bytes(model[:arg]) and isinstance(got, bytes)
                del model[:arg]
            assert len(ring) == len(model)
        assert ring.read(1 << 20) == bytes(model)

    check()


def test_ticker_survives_a_new_event_loop():
    """asyncio.run per request (the reference's scripts do that): the ticker re-binds to the running loop."""
    t = tk.DecodeTicker(lambda ws, keys: [fake_convert(w) for w in ws], in_thread=False)
    w = [5] * 28
    for _ in range(3):
        assert asyncio.run(t.decode(w, 1)) == fake_convert(w)
    assert t.ticks == 3


def test_adapters_over_gpu_ring_slots(speechpipe):
    """N3: with gpu_ring the decoder yields byte counts and the PCM waits in a ring slot (here a host stand-in): same
    bytes as the per-stream path for even AND odd pull sizes, ticks not windows, ring-bound and plain windows share a
    tick, slots are returned at eos and at reset."""
    n = 40
    streams = [sp.synth_token_strings(500 + i, 4 + (i % 7)) for i in range(n)]
    want = [b"".join(sp.decode_stream(s, fake_convert)) for s in streams]

    async def pull_loop(ad, size):
        out = bytearray()
        while True:
            c = await ad.pull(size)
            assert len(c.pcm) == size or c.eos or not c.pcm
            out += c.pcm
            if c.eos:
                return bytes(out)

    async def main():
        ads = [SnacB200Adapter("p", token_source=_source(s, gap=5), seed=i, gpu_ring=(i % 4 != 3)) for i, s in enumerate(streams)]
        return await asyncio.gather(*[pull_loop(a, (333, 64, 4096, 1001)[i % 4]) for i, a in enumerate(ads)])

    got = asyncio.run(main())
    assert got == want
    st = speechpipe.get_ticker().stats()
    assert len(speechpipe.calls) >= st["ticks"] and st["ticks"] * 6 < st["windows"], st   # <= 2 engine calls per mixed tick
    assert len(speechpipe.calls) <= 2 * st["ticks"]
    assert len(speechpipe.fake_ring._free) == 128 and all(not b for b in speechpipe.fake_ring.bufs)

    async def barge():
        ad = SnacB200Adapter("p", token_source=_source(streams[0], gap=1), seed=0, gpu_ring=True)
        first = await ad.pull(100)
        assert len(speechpipe.fake_ring._free) == 127
        await ad.reset()
        assert len(speechpipe.fake_ring._free) == 128
        return first.pcm, await pull_loop(ad, 777)

    first, again = asyncio.run(barge())
    assert first == want[0][:100] and again == want[0]
    assert len(speechpipe.fake_ring._free) == 128


def test_more_requests_than_ring_slots_fall_back_to_the_host_ring(speechpipe, monkeypatch):
    """Auto mode (gpu_ring=None on a CUDA host): when every ring slot is taken a request keeps its PCM in host bytes
    instead of failing; asking for the GPU ring explicitly raises."""
    from project_morpheus_b200 import _lib

    monkeypatch.setattr(speechpipe, "snac_device", "cuda")
    monkeypatch.setenv("SNACB_GPU_RING", "1")
    small = FakeGpuRing(2)
    monkeypatch.setattr(speechpipe, "get_ring", lambda: small)
    speechpipe.fake_ring.bufs = small.bufs  # the stand-in batch function writes into fake_ring.bufs
    streams = [sp.synth_token_strings(700 + i, 5) for i in range(4)]
    want = [b"".join(sp.decode_stream(s, fake_convert)) for s in streams]

    async def drain(ad):
        out = bytearray()
        while True:
            c = await ad.pull(512)
            out += c.pcm
            if c.eos:
                return bytes(out)

    async def main():
        ads = [SnacB200Adapter("p", token_source=_source(s, gap=3), seed=i) for i, s in enumerate(streams)]
        res = await asyncio.gather(*[drain(a) for a in ads])
        return res

    assert asyncio.run(main()) == want
    assert len(small._free) == 2

    async def explicit():
        hold = [SnacB200Adapter("p", token_source=_source(streams[0], gap=1), gpu_ring=True) for _ in range(3)]
        firsts = [await hold[0].pull(8), await hold[1].pull(8)]
        with pytest.raises(_lib.SnacbError):
            await hold[2].pull(8)
        for h in hold[:2]:
            await h.reset()
        return firsts

    asyncio.run(explicit())
    assert len(small._free) == 2
