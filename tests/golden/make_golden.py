"""Regenerates ``tests/golden/speechpipe_golden.json`` by running the VERBATIM reference file.

Run in the authoring container only (needs ``/root/reference``):

    python tests/golden/make_golden.py

The reference's own tests hold no golden vector for this path, so the fixtures are
outputs of the reference ``Morpheus_Client/tts_engine/speechpipe.py`` itself, executed unmodified
with ``oracle.snac_ref`` injected as the ``snac`` module (the third-party package is absent).
What is recorded:

  G1  de-interleave + validator KATs: the code tensors the reference hands to ``model.decode`` and
      whether ``convert_to_audio`` returned None                       (speechpipe.py:72-111)
  G2  ``turn_token_into_id`` table                                     (speechpipe.py:146-189)
  G3  ``tokens_decoder`` chunk-size sequences                          (speechpipe.py:191-293)
  G4  PCM of the config-1 stream (14 frames) with injected noise, W1 weights regenerated from
      seed (the 52 MB of weights are not committed); stored as int16 lists for three chunks plus
      a sha256 of every chunk.
"""
from __future__ import annotations

import asyncio
import hashlib
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_loader, snac_ref, speechpipe_ref as sp  # noqa: E402
from project_morpheus_b200 import weights  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "speechpipe_golden.json")

KAT_WINDOWS = [
    list(range(1, 8)),
    list(range(100, 128)),
    [4095] * 28,
    [0] * 28,
    [4096] + [5] * 27,
    [5] * 27 + [4097],
    [-1] + [7] * 27,
    [3] * 6,
    [11, 22, 33, 44, 55, 66, 77, 88, 99],
    list(range(1, 50)),
    [int(x) for x in sp.synth_codes(3, 4)],
    [int(x) for x in sp.synth_codes(4, 7)],
    [int(x) for x in sp.synth_codes(5, 5)] + [9, 9, 9],
]

TOKEN_STRINGS = [
    ("<custom_token_10>", 0), ("<custom_token_11>", 0), ("<custom_token_4106>", 1), ("<custom_token_4105>", 1),
    ("<custom_token_28681>", 6), ("<custom_token_28682>", 13), ("<custom_token_5>", 0), ("<custom_token_123456>", 3),
    ("junk<custom_token_20><custom_token_30>", 0), ("  <custom_token_77>  ", 2), ("<custom_token_77", 2),
    ("<custom_token_abc>", 0), ("hello", 0), ("", 5), ("<custom_token_>", 0), ("<custom_token_12>x", 0),
    ("<custom_token_-5>", 0), ("<custom_token_ 15>", 0),
]


def main() -> None:
    torch.set_grad_enabled(False)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    sd = weights.random_state_dict(0, "w1")
    ref = ref_loader.load_reference_speechpipe(sd)
    model = ref.model
    gold = {"generator": "tests/golden/make_golden.py", "reference": "Morpheus_Client/tts_engine/speechpipe.py",
            "weights": "random_state_dict(0,'w1')", "torch": torch.__version__}

    # ---- G1
    seen = {}
    real_decode = model.decode

    def spy(codes):
        seen["codes"] = [c.clone() for c in codes]
        raise _Stop()

    class _Stop(Exception):
        pass

    g1 = []
    model.decode = spy
    for win in KAT_WINDOWS:
        seen.clear()
        try:
            out = ref.convert_to_audio(list(win), 0)
            verdict = "none" if out is None else "bytes"
        except _Stop:
            verdict = "decode"
        row = {"tokens": list(win), "verdict": verdict}
        if "codes" in seen:
            row["codes"] = [c.reshape(-1).tolist() for c in seen["codes"]]
        g1.append(row)
    model.decode = real_decode
    gold["g1_deinterleave"] = g1

    # ---- G2
    gold["g2_token_ids"] = [{"text": t, "index": i, "id": ref.turn_token_into_id(t, i)} for t, i in TOKEN_STRINGS]

    # ---- G3 (noise off: only sizes matter)
    model.set_noise("off")

    async def run_stream(strings):
        async def gen():
            for s in strings:
                yield s
        return [c async for c in ref.tokens_decoder(gen())]

    g3 = []
    for frames in (1, 3, 4, 5, 8, 10):
        chunks = asyncio.run(run_stream(sp.synth_token_strings(frames, frames)))
        g3.append({"frames": frames, "sizes": [len(c) for c in chunks]})
    # a stream with dropped tokens (id 0, negatives, junk) interleaved: slot indices shift (Q2)
    dirty = sp.synth_token_strings(7, 6)
    dirty.insert(3, "<custom_token_10>")  # id 0 at slot 3 -> dropped
    dirty.insert(9, "noise")
    dirty.insert(20, "<custom_token_3>")  # negative id
    chunks = asyncio.run(run_stream(dirty))
    g3.append({"frames": "dirty6", "sizes": [len(c) for c in chunks],
               "sha256": [hashlib.sha256(c).hexdigest() for c in chunks]})
    gold["g3_chunk_sizes"] = g3

    # ---- G4: config-1 stream, noise injected per decode call
    frames = 14
    strings = sp.synth_token_strings(0, frames)
    call = {"n": 0}

    def decode_with_noise(codes):
        F = codes[0].shape[1]
        model.set_noise(snac_ref.make_noise(1, F, seed=99 + call["n"]))
        call["n"] += 1
        return real_decode(codes)

    model.decode = decode_with_noise
    chunks = asyncio.run(run_stream(strings))
    model.decode = real_decode
    keep = (1, 5, len(chunks) - 1)
    gold["g4_config1"] = {
        "frames": frames, "stream": 0, "noise": "make_noise(1, F, seed=99+call_index)",
        "sizes": [len(c) for c in chunks],
        "sha256": [hashlib.sha256(c).hexdigest() for c in chunks],
        "pcm": {str(i): np.frombuffer(chunks[i], dtype="<i2").tolist() for i in keep},
        "rms": [float(np.sqrt(np.mean(np.frombuffer(c, dtype="<i2").astype(np.float64) ** 2))) if c else 0.0 for c in chunks],
    }

    with open(OUT, "w") as f:
        json.dump(gold, f, separators=(",", ":"))
    print(OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
