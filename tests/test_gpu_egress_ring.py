"""N3 on the GPU (csrc/egress_ring.cu): decode ticks written straight into pinned per-stream PCM rings.  B200 only.

Byte-exactness ladder: the ring kernel == the golden vectors the verbatim reference stitcher produced
(tests/golden/egress_golden.json) == the native host stitcher == the numpy oracle on random streams; then the decode
tick into the rings == the decode tick into a host matrix; then adapters over the rings == the per-stream decoder."""
import hashlib
import json
import os

import numpy as np
import pytest
import torch

from oracle import egress_ref
from project_morpheus_b200 import _lib, egress

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "egress_golden.json")


def push_chunks(ring, slot, chunks, eos_last):
    """One stream's chunks through the ring kernel, one push per chunk; returns every byte the slot emitted."""
    out = bytearray()
    for i, c in enumerate(chunks):
        a = np.frombuffer(c, dtype="<i2")
        d = torch.from_numpy(a.copy()).cuda() if len(a) else torch.zeros(1, dtype=torch.int16, device="cuda")
        ring.push_device([slot], d.data_ptr(), max(1, len(a)), len(a), None, eos=[1 if (eos_last and i == len(chunks) - 1) else 0])
        ring.sync()
        out += ring.read(slot, 1 << 20)
    if not eos_last:
        ring.flush(slot)
        out += ring.read(slot, 1 << 20)
    return bytes(out)


def test_ring_kernel_matches_golden_vectors_of_the_reference_stitcher(ensure_lib):
    doc = json.load(open(GOLD))
    rng = np.random.default_rng(20240607)
    for row in doc["stitch"]:
        chunks = [rng.integers(-32768, 32768, size=n).astype("<i2").tobytes() for n in row["sizes"]]
        assert [hashlib.sha256(c).hexdigest() for c in chunks] == row["chunks_sha256"]
        ring = egress.GpuPcmRing(2, 8192, overlap_ms=row["overlap_ms"])
        got = push_chunks(ring, 1, chunks, row["eos_last"])
        # the golden rows hold the sha of every yielded chunk; the ring is the byte stream of those chunks in order
        orc = b"".join(p for p, _ in egress_ref.stitch([(c, row["eos_last"] and i == len(chunks) - 1) for i, c in enumerate(chunks)],
                                                      24000, row["overlap_ms"]))
        assert [hashlib.sha256(p).hexdigest() for p, _ in egress_ref.stitch(
            [(c, row["eos_last"] and i == len(chunks) - 1) for i, c in enumerate(chunks)], 24000, row["overlap_ms"])] == [w["sha256"] for w in row["out"]]
        assert got == orc, row["seed_index"]
        ring.close()


@pytest.mark.parametrize("overlap_ms", [0.0, 0.04, 5.5, 10.0, 40.0])
def test_ring_kernel_equals_host_stitcher_on_random_ticks(ensure_lib, overlap_ms):
    """Many slots per tick, windows missing (status != OK), streams ending with and without eos, odd write positions,
    ring wrap-around (small rings): every slot's byte stream equals the native host stitcher's."""
    rng = np.random.default_rng(int(overlap_ms * 100) + 3)
    n_slots, length = 37, 2048
    ring = egress.GpuPcmRing(n_slots, 8192, overlap_ms=overlap_ms)
    host = [egress.Stitcher(24000, overlap_ms) for _ in range(n_slots)]
    want = [bytearray() for _ in range(n_slots)]
    got = [bytearray() for _ in range(n_slots)]
    done = [False] * n_slots
    for tick in range(14):
        live = [s for s in range(n_slots) if not done[s] and rng.random() < 0.8]
        if not live:
            continue
        rng.shuffle(live)
        n = len(live)
        pcm = rng.integers(-32768, 32768, size=(n, length + 8)).astype(np.int16)  # row stride > len
        status = rng.choice([_lib.WIN_OK, _lib.WIN_OK, _lib.WIN_OK, _lib.WIN_REJECTED, _lib.WIN_EMPTY], size=n).astype(np.int32)
        eos = (rng.random(n) < 0.07).astype(np.int32)
        for s in live:  # a consumer that fell behind reads before its ring would refuse the tick
            if ring.room(s) < length:
                got[s] += ring.read(s, 1 << 20)
        d_pcm, d_st = torch.from_numpy(pcm).cuda(), torch.from_numpy(status).cuda()
        ring.push_device(live, d_pcm.data_ptr(), length + 8, length, d_st.data_ptr(), eos=eos)
        ring.sync()
        for i, s in enumerate(live):
            chunk = pcm[i, :length].tobytes() if status[i] == _lib.WIN_OK else b""
            if chunk or eos[i]:
                data, _ = host[s].push(chunk, bool(eos[i]))
                if data:
                    want[s] += data
            done[s] = done[s] or bool(eos[i])
            got[s] += ring.read(s, 2 * int(rng.integers(1, 3000)))  # partial reads: the rest stays in the ring
    for s in range(n_slots):
        if not done[s]:
            ring.flush(s)
            tail = host[s].flush()
            if tail:
                want[s] += tail
        got[s] += ring.read(s, 1 << 20)
        assert bytes(got[s]) == bytes(want[s]), (s, len(got[s]), len(want[s]))
        host[s].close()
    assert sum(len(w) for w in want) > 100 * 4096
    ring.close()


def test_ring_refuses_to_overflow_and_rejects_bad_ticks(ensure_lib):
    ring = egress.GpuPcmRing(4, 4096, overlap_ms=0.0)
    pcm = torch.arange(2 * 2048, dtype=torch.int16, device="cuda").reshape(2, 2048)
    ring.push_device([0, 1], pcm.data_ptr(), 2048, 2048)
    ring.push_device([0, 3], pcm.data_ptr(), 2048, 2048)
    assert ring.room(0) == 0 and ring.room(2) == 4096
    with pytest.raises(_lib.SnacbError, match="full"):
        ring.push_device([0], pcm.data_ptr(), 2048, 2048)
    with pytest.raises(_lib.SnacbError, match="twice"):
        ring.push_device([2, 2], pcm.data_ptr(), 2048, 2048)
    with pytest.raises(_lib.SnacbError, match="out of range"):
        ring.push_device([4], pcm.data_ptr(), 2048, 2048)
    ring.sync()
    assert ring.available(0) == 2 * 4096 and ring.available(1) == 4096 and ring.available(2) == 0
    a = np.frombuffer(ring.read(0, 8192), dtype="<i2")
    assert np.array_equal(a[:2048], pcm[0].cpu().numpy()) and np.array_equal(a[2048:], pcm[0].cpu().numpy())
    assert ring.room(0) == 4096
    ring.push_device([0], pcm[1:].data_ptr(), 2048, 2048)      # wraps onto the start of the ring
    ring.sync()
    assert np.array_equal(np.frombuffer(ring.read(0, 4096), dtype="<i2"), pcm[1].cpu().numpy())
    ring.reset(1)
    assert ring.available(1) == 0 and ring.read(1, 64) == b""
    ring.push_device([1], pcm.data_ptr(), 2048, 2048)          # the first push after a reset starts the slot from scratch
    ring.sync()
    assert np.array_equal(np.frombuffer(ring.read(1, 1 << 20), dtype="<i2"), pcm[0].cpu().numpy())
    with pytest.raises(_lib.SnacbError):
        egress.GpuPcmRing(4, 4100)           # not a multiple of 8
    with pytest.raises(_lib.SnacbError):
        egress.GpuPcmRing(4, 4096, overlap_ms=200.0)   # 2 * overlap + 2048 > ring
    ring.close()


@pytest.mark.parametrize("precision", ["fp16", "fp32"])
def test_decode_tick_into_rings_equals_decode_tick_into_host_matrix(state_dict_w1, ensure_lib, precision):
    """snacb_decode_windows_to_ring == snacb_decode_windows_host + concatenation per stream: ragged windows (rejected,
    single-frame, 4- and 7-frame), several ticks, shuffled slot order, wrap-around; statuses identical."""
    from helpers import windows_tokens
    from project_morpheus_b200.engine import SnacEngine

    eng = SnacEngine(state_dict_w1, device=0, precision=precision)
    n = 48
    ring = egress.GpuPcmRing(n, 8192)
    rng = np.random.default_rng(11)
    want = [bytearray() for _ in range(n)]
    got = [bytearray() for _ in range(n)]
    for tick in range(7):
        slots = rng.permutation(n)[: int(rng.integers(8, n + 1))].astype(np.int32)
        lens = rng.choice([0, 5, 7, 28, 28, 28, 49], size=len(slots))
        tok = np.zeros((len(slots), 49), dtype=np.int32)
        for i, ln in enumerate(lens):
            if ln:
                tok[i, :ln] = windows_tokens(1, 7, 1000 * tick + i)[0, :ln]
        if tick == 3:
            tok[0, 3] = 4096 + 7 * 0   # a poisoned window (code 4096): no PCM for that slot, everybody else unaffected
        keys = rng.integers(0, 2**63, size=len(slots), dtype=np.uint64)
        pcm, st = eng.decode_windows(tok, ntok=lens, noise="philox", seed=5, keys=keys)
        pcm, st = pcm.copy(), st.copy()
        st2, emitted = eng.decode_windows_to_ring(ring, tok, slots, ntok=lens, noise="philox", seed=5, keys=keys)
        assert np.array_equal(st, st2) and np.array_equal(emitted, np.where(st == _lib.WIN_OK, 2048, 0))
        for i, s in enumerate(slots):
            if st[i] == _lib.WIN_OK:
                want[s] += pcm[i].tobytes()
            got[s] += ring.read(int(s), 1 << 20)
    assert all(bytes(g) == bytes(w) for g, w in zip(got, want))
    assert sum(len(w) for w in want) > 50 * 4096
    ring.close()
    eng.close()


def test_stalled_consumer_loses_its_window_not_the_tick(state_dict_w1, ensure_lib):
    from helpers import windows_tokens
    from project_morpheus_b200.engine import SnacEngine

    eng = SnacEngine(state_dict_w1, device=0, precision="fp16")
    ring = egress.GpuPcmRing(3, 4096)
    tok = windows_tokens(2, 4, 77)
    for tick in range(3):
        st, em = eng.decode_windows_to_ring(ring, tok, [0, 2], noise="off")
        assert list(st) == [_lib.WIN_OK, _lib.WIN_OK]
        assert list(em) == ([2048, 2048] if tick < 2 else [-1, 2048])   # slot 0 is never read: full after two windows
        assert len(ring.read(2, 1 << 20)) == 4096
    assert ring.available(0) == 2 * 4096 and ring.room(0) == 0
    first = ring.read(0, 4096)
    pcm, _ = eng.decode_windows(tok, noise="off")
    assert first == pcm[0].tobytes()
    ring.close()
    eng.close()


def test_adapters_with_gpu_crossfade_equal_host_stitcher_over_the_serial_decoder(state_dict_w1, monkeypatch):
    """The integrated path with a non-zero overlap: adapters pull from GPU rings that crossfade consecutive windows of a
    stream (SNACB_RING_OVERLAP_MS = 10) == the native host stitcher applied to the chunks of the serial per-stream
    decoder (noise off), including the tail flushed at the end of each stream."""
    import asyncio
    import importlib
    import sys

    from oracle import speechpipe_ref as sp

    monkeypatch.setenv("SNACB_NOISE", "off")
    monkeypatch.setenv("SNACB_PRECISION", "fp16")
    monkeypatch.setenv("SNACB_RANDOM_INIT", "0:w1")
    monkeypatch.setenv("SNACB_GPU_RING", "1")
    monkeypatch.setenv("SNACB_RING_SLOTS", "32")
    monkeypatch.setenv("SNACB_RING_OVERLAP_MS", "10")
    monkeypatch.delenv("ORPHEUS_SNAC_PATH", raising=False)
    sys.modules.pop("project_morpheus_b200.speechpipe", None)
    speechpipe = importlib.import_module("project_morpheus_b200.speechpipe")
    from project_morpheus_b200.adapter import SnacB200Adapter

    n = 24
    streams = [sp.synth_token_strings(900 + i, 3 + (i % 6)) for i in range(n)]

    def source(strings):
        async def gen(**_):
            for i, s in enumerate(strings):
                if i % 5 == 0:
                    await asyncio.sleep(0)
                yield s
        return gen

    async def pull_loop(ad, size):
        out = bytearray()
        while True:
            c = await ad.pull(size)
            out += c.pcm
            if c.eos:
                return bytes(out)

    async def main():
        ads = [SnacB200Adapter("p", token_source=source(s), seed=i) for i, s in enumerate(streams)]
        return await asyncio.gather(*[pull_loop(a, (4096, 1000, 333)[i % 3]) for i, a in enumerate(ads)])

    got = asyncio.run(main())
    assert speechpipe.get_ring().overlap_samples == 240 and len(speechpipe.get_ring()._free) == 32

    async def serial(strings):
        return [c async for c in speechpipe.tokens_decoder(source(strings)(), ticker=False)]

    for i, s in enumerate(streams):
        st = egress.Stitcher(24000, 10.0)
        want = bytearray()
        for chunk in asyncio.run(serial(s)):
            if chunk:
                data, _ = st.push(chunk, False)
                if data:
                    want += data
        want += st.flush() or b""
        st.close()
        assert got[i] == bytes(want), (i, len(got[i]), len(want))
        assert len(want) >= 4096
    assert sum(len(g) > 3 * 4096 for g in got) >= n // 2   # most streams crossfade several windows
    sys.modules.pop("project_morpheus_b200.speechpipe", None)


def test_barge_in_storm_over_gpu_rings(state_dict_w1, monkeypatch):
    """Slot reuse under fire: 96 adapters on 40 ring slots' worth of concurrency, a third of them reset (barge-in) at a
    random point while their window may be queued or on the GPU, slots handed straight to new requests.  Every request
    that runs to its end delivers exactly the bytes of the serial decoder, and every slot comes back."""
    import asyncio
    import importlib
    import random
    import sys

    from oracle import speechpipe_ref as sp

    monkeypatch.setenv("SNACB_NOISE", "off")
    monkeypatch.setenv("SNACB_PRECISION", "fp16")
    monkeypatch.setenv("SNACB_RANDOM_INIT", "0:w1")
    monkeypatch.setenv("SNACB_GPU_RING", "1")
    monkeypatch.setenv("SNACB_RING_SLOTS", "40")
    monkeypatch.setenv("SNACB_RING_OVERLAP_MS", "0")
    monkeypatch.delenv("ORPHEUS_SNAC_PATH", raising=False)
    sys.modules.pop("project_morpheus_b200.speechpipe", None)
    speechpipe = importlib.import_module("project_morpheus_b200.speechpipe")
    from project_morpheus_b200.adapter import SnacB200Adapter

    rnd = random.Random(7)
    n = 96
    streams = [sp.synth_token_strings(1200 + i, 4 + (i % 8)) for i in range(n)]

    def source(strings):
        async def gen(**_):
            for i, s in enumerate(strings):
                if i % 3 == 0:
                    await asyncio.sleep(0)
                yield s
        return gen

    async def request(i, gate):
        async with gate:  # at most 40 requests alive: one ring slot each
            ad = SnacB200Adapter("p", token_source=source(streams[i]), seed=i, gpu_ring=True)
            size = rnd.choice([256, 1000, 4096])
            if i % 3 == 0:  # barge-in: pull a little, reset, then play the request from the start
                for _ in range(rnd.randint(0, 6)):
                    if (await ad.pull(size)).eos:
                        break
                await ad.reset()
            out = bytearray()
            while True:
                c = await ad.pull(size)
                out += c.pcm
                if c.eos:
                    return bytes(out)

    async def main():
        gate = asyncio.Semaphore(40)
        return await asyncio.gather(*[request(i, gate) for i in range(n)])

    got = asyncio.run(main())
    assert len(speechpipe.get_ring()._free) == 40

    async def serial(strings):
        return b"".join([c async for c in speechpipe.tokens_decoder(source(strings)(), ticker=False)])

    for i in range(n):
        assert got[i] == asyncio.run(serial(streams[i])), i
    sys.modules.pop("project_morpheus_b200.speechpipe", None)
