"""Generates tests/golden/egress_golden.json by running the VERBATIM reference stitcher / riff_header
(/root/reference/Morpheus_Client/orchestrator/stitcher.py, server.py) on seeded chunk sequences.  Run in the authoring
container only (the reference tree is not on the GPU box): python tests/golden/make_golden_egress.py"""
import asyncio
import hashlib
import importlib
import json
import os
import re
import sys
import types

import numpy as np

REF = os.environ.get("MORPHEUS_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "egress_golden.json")


def load_reference_stitcher():
    pkg = types.ModuleType("_mref_orch")
    pkg.__path__ = [os.path.join(REF, "Morpheus_Client", "orchestrator")]
    sys.modules["_mref_orch"] = pkg
    return importlib.import_module("_mref_orch.stitcher"), importlib.import_module("_mref_orch.adapter")


def reference_riff_header(sample_rate):
    """server.py imports the whole web stack; its riff_header is a pure struct.pack: evaluate that function alone."""
    src = open(os.path.join(REF, "Morpheus_Client", "server.py")).read()
    m = re.search(r"def riff_header\(.*?\n(?=\n\nasync def )", src, flags=re.S)
    ns = {"struct": __import__("struct"), "SAMPLE_RATE": 24000}
    exec(m.group(0), ns)
    return ns["riff_header"](sample_rate)


def cases():
    rng = np.random.default_rng(20240607)
    out = []
    for idx, (overlap_ms, sizes, eos_last) in enumerate([
        (0.0, [2048, 2048, 2048], True), (0.0, [2048, 0, 100], False), (10.0, [2048, 2048, 2048, 2048], True),
        (10.0, [2048, 2048, 2048], False), (10.0, [100, 100, 100, 2048, 50, 2048], True), (25.0, [2048] * 6, False),
        (5.5, [300, 7, 2048, 1, 2048], True), (100.0, [2048, 2048], False), (10.0, [], False), (10.0, [0, 0, 500], True),
        (0.04, [64, 64, 64], True), (10.0, [240, 240, 241, 239, 2048], False),
    ]):
        chunks = [rng.integers(-32768, 32768, size=n).astype("<i2").tobytes() for n in sizes]
        out.append({"overlap_ms": overlap_ms, "sizes": sizes, "eos_last": eos_last, "seed_index": idx,
                    "chunks_hex": [c.hex() for c in chunks]})
    return out


def main():
    stitcher, adapter = load_reference_stitcher()

    async def run(case):
        async def gen():
            n = len(case["chunks_hex"])
            for i, h in enumerate(case["chunks_hex"]):
                yield adapter.AudioChunk(pcm=bytes.fromhex(h), duration_ms=0.0, markers={"i": i}, eos=case["eos_last"] and i == n - 1)
        return [c async for c in stitcher.stitch_chunks(gen(), sample_rate=24000, overlap_ms=case["overlap_ms"], emit_markers=True)]

    rows = []
    for case in cases():
        got = asyncio.run(run(case))
        rows.append({**{k: case[k] for k in ("overlap_ms", "sizes", "eos_last", "seed_index")},
                     "chunks_sha256": [hashlib.sha256(bytes.fromhex(h)).hexdigest() for h in case["chunks_hex"]],
                     "out": [{"n": len(c.pcm) // 2, "eos": bool(c.eos), "sha256": hashlib.sha256(c.pcm).hexdigest(),
                              "duration_ms": c.duration_ms, "marker": c.markers} for c in got]})
    doc = {"generator": "tests/golden/make_golden_egress.py", "reference": "Morpheus_Client/orchestrator/stitcher.py, server.py:riff_header",
           "rng": "numpy default_rng(20240607), chunk i = integers(-32768, 32768, size).astype('<i2') in case order",
           "riff_header_24000_hex": reference_riff_header(24000).hex(), "riff_header_16000_hex": reference_riff_header(16000).hex(),
           "stitch": rows}
    json.dump(doc, open(OUT, "w"), indent=1)
    print("wrote", OUT, len(rows), "cases")


if __name__ == "__main__":
    main()
