"""N3 (PCM egress) without a GPU: RIFF header and overlap-add stitcher - native (csrc/egress.cpp) == numpy oracle ==
golden vectors produced by the verbatim reference (and the verbatim reference itself where its tree is mounted)."""
import asyncio
import hashlib
import json
import os

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from oracle import egress_ref
from project_morpheus_b200 import egress
from project_morpheus_b200.adapter import AudioChunk

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "egress_golden.json")


def gold():
    return json.load(open(GOLD))


def regen_chunks(doc):
    """The generating script's chunk bytes, re-drawn from its documented RNG (checked against their hashes)."""
    rng = np.random.default_rng(20240607)
    out = []
    for row in doc["stitch"]:
        chunks = [rng.integers(-32768, 32768, size=n).astype("<i2").tobytes() for n in row["sizes"]]
        assert [hashlib.sha256(c).hexdigest() for c in chunks] == row["chunks_sha256"]
        out.append(chunks)
    return out


def run_native(chunks, eos_last, overlap_ms, markers=True):
    async def go():
        async def gen():
            for i, c in enumerate(chunks):
                yield AudioChunk(pcm=c, duration_ms=0.0, markers={"i": i}, eos=eos_last and i == len(chunks) - 1)
        return [c async for c in egress.stitch_chunks(gen(), sample_rate=24000, overlap_ms=overlap_ms, emit_markers=markers)]
    return asyncio.run(go())


def test_riff_header_matches_reference_bytes(ensure_lib):
    doc = gold()
    assert egress.riff_header(24000).hex() == doc["riff_header_24000_hex"] == egress_ref.riff_header(24000).hex()
    assert egress.riff_header(16000).hex() == doc["riff_header_16000_hex"] == egress_ref.riff_header(16000).hex()
    assert len(egress.riff_header()) == 44


def test_stitcher_matches_golden_vectors_of_the_reference(ensure_lib):
    doc = gold()
    for row, chunks in zip(doc["stitch"], regen_chunks(doc)):
        want = row["out"]
        got = run_native(chunks, row["eos_last"], row["overlap_ms"])
        orc = list(egress_ref.stitch([(c, row["eos_last"] and i == len(chunks) - 1) for i, c in enumerate(chunks)], 24000, row["overlap_ms"]))
        assert [(len(c.pcm) // 2, c.eos, hashlib.sha256(c.pcm).hexdigest()) for c in got] == [(w["n"], w["eos"], w["sha256"]) for w in want], row["seed_index"]
        assert [(hashlib.sha256(p).hexdigest(), e) for p, e in orc] == [(w["sha256"], w["eos"]) for w in want], row["seed_index"]
        assert [c.duration_ms for c in got] == [w["duration_ms"] for w in want]
        assert [c.markers for c in got] == [w["marker"] for w in want]
    assert all(c.markers is None for c in run_native(regen_chunks(doc)[2], True, 10.0, markers=False))


@settings(max_examples=150, deadline=None)
@given(st.lists(st.integers(min_value=0, max_value=700), min_size=0, max_size=7), st.sampled_from([0.0, 0.04, 1.0, 5.5, 10.0, 25.0]),
       st.booleans(), st.integers(min_value=0, max_value=2**31 - 1))
def test_stitcher_equals_numpy_oracle_on_random_streams(sizes, overlap_ms, eos_last, seed):
    rng = np.random.default_rng(seed)
    chunks = [rng.integers(-32768, 32768, size=n).astype("<i2").tobytes() for n in sizes]
    got = run_native(chunks, eos_last, overlap_ms)
    want = list(egress_ref.stitch([(c, eos_last and i == len(chunks) - 1) for i, c in enumerate(chunks)], 24000, overlap_ms))
    assert [(c.pcm, c.eos) for c in got] == want


def test_oracle_matches_verbatim_reference():
    ref_root = os.environ.get("MORPHEUS_REFERENCE_ROOT", "/root/reference")
    if not os.path.isfile(os.path.join(ref_root, "Morpheus_Client", "orchestrator", "stitcher.py")):
        pytest.skip("reference tree not mounted")
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden_egress", os.path.join(os.path.dirname(GOLD), "make_golden_egress.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    stitcher, adapter = mod.load_reference_stitcher()
    rng = np.random.default_rng(5)
    for overlap_ms in (0.0, 3.0, 10.0, 40.0):
        for trial in range(6):
            sizes = rng.integers(0, 900, size=int(rng.integers(0, 7))).tolist()
            eos_last = bool(rng.integers(0, 2))
            chunks = [rng.integers(-32768, 32768, size=n).astype("<i2").tobytes() for n in sizes]

            async def run():
                async def gen():
                    for i, c in enumerate(chunks):
                        yield adapter.AudioChunk(pcm=c, duration_ms=0.0, eos=eos_last and i == len(chunks) - 1)
                return [c async for c in stitcher.stitch_chunks(gen(), sample_rate=24000, overlap_ms=overlap_ms)]

            ref = [(c.pcm, c.eos) for c in asyncio.run(run())]
            assert list(egress_ref.stitch([(c, eos_last and i == len(chunks) - 1) for i, c in enumerate(chunks)], 24000, overlap_ms)) == ref
            assert [(c.pcm, c.eos) for c in run_native(chunks, eos_last, overlap_ms)] == ref


def test_stitch_argument_checks(ensure_lib):
    s = egress.Stitcher(24000, 10.0)
    assert s.overlap_samples == 240
    assert s.push(b"", False) == (None, False)
    out, eos = s.push(np.arange(300, dtype="<i2").tobytes(), False)
    assert len(out) == 2 * 60 and not eos
    assert len(s.flush()) == 2 * 240 and s.flush() is None
    s.close()


def test_stitcher_bank_tick_equals_per_stream_stitchers(ensure_lib):
    rng = np.random.default_rng(11)
    n, ticks = 37, 6
    bank = egress.StitcherBank(n, 24000, 10.0)
    singles = [egress.Stitcher(24000, 10.0) for _ in range(n)]
    for t in range(ticks):
        pcm = rng.integers(-32768, 32768, size=(n, 2048)).astype(np.int16)
        eos = (rng.random(n) < 0.1).astype(np.int32) if t == ticks - 1 else None
        out, out_len, out_eos = bank.push_tick(np.arange(n), pcm, eos)
        for i in range(n):
            data, e = singles[i].push(pcm[i].tobytes(), bool(eos[i]) if eos is not None else False)
            if data is None:
                assert out_len[i] == -1
            else:
                assert out[i, : out_len[i]].tobytes() == data and bool(out_eos[i]) == e
    bank.reset(3)
    out, out_len, _ = bank.push_tick([3], np.zeros((1, 2048), dtype=np.int16))
    assert out_len[0] == 2048 - 240
    bank.close()
