"""The independent restatement (oracle/speechpipe_ref.py) and the product's host logic against the
VERBATIM reference file, on seeded and adversarial streams.  Authoring container only."""
import asyncio

import numpy as np
import pytest
import torch

from oracle import ref_loader, snac_ref, speechpipe_ref as sp

pytestmark = pytest.mark.needs_reference
if not ref_loader.reference_available():
    pytest.skip("/root/reference is not mounted", allow_module_level=True)


@pytest.fixture(scope="module")
def ref(state_dict_w1):
    torch.set_grad_enabled(False)
    mod = ref_loader.load_reference_speechpipe(state_dict_w1, quiet=True)
    mod.model.set_noise("off")
    return mod


def _run_ref(ref, strings):
    async def go():
        async def gen():
            for s in strings:
                yield s
        return [c async for c in ref.tokens_decoder(gen())]
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        return asyncio.run(go())


def _oracle_convert(model):
    def decode(c0, c1, c2):
        codes = [torch.from_numpy(c.astype(np.int64))[None] for c in (c0, c1, c2)]
        return model.decode(codes)[0, 0].numpy()
    return lambda win: sp.window_to_pcm(win, decode)


@pytest.mark.parametrize("frames", [1, 2, 3, 4, 5, 8, 11])
def test_stream_bytes_identical(ref, frames):
    strings = sp.synth_token_strings(40 + frames, frames)
    want = _run_ref(ref, strings)
    got = list(sp.decode_stream(strings, _oracle_convert(ref.model)))
    assert [len(c) for c in got] == [len(c) for c in want]
    assert got == want


def test_dirty_stream_bytes_identical(ref):
    rng = np.random.default_rng(7)
    strings = sp.synth_token_strings(9, 9)
    for pos in sorted(rng.integers(0, len(strings), 12).tolist(), reverse=True):
        strings.insert(pos, rng.choice(["<custom_token_10>", "zzz", "<custom_token_2>", "<custom_token_99", ""]))
    strings[5] = strings[5] + strings[6]  # two tokens in one string: the last one wins
    want = _run_ref(ref, strings)
    got = list(sp.decode_stream(strings, _oracle_convert(ref.model)))
    assert got == want


def test_convert_to_audio_edges(ref):
    conv = _oracle_convert(ref.model)
    for win in ([5] * 6, [5] * 7, [5] * 13, [-1] + [5] * 27, [4097] + [5] * 27, [0] * 28, list(range(1, 29))):
        import contextlib, io
        with contextlib.redirect_stdout(io.StringIO()):
            want = ref.convert_to_audio(list(win), 0)
        assert conv(list(win)) == want
    with pytest.raises(IndexError):
        ref.convert_to_audio([4096] + [5] * 27, 0)
    with pytest.raises(IndexError):
        conv([4096] + [5] * 27)


def test_turn_token_into_id_matches(ref):
    from project_morpheus_b200 import tokens
    rng = np.random.default_rng(3)
    texts = [f"<custom_token_{int(n)}>" for n in rng.integers(0, 7 * 4096 + 20, 200)]
    texts += ["", "x", "<custom_token_", "<custom_token_1>>", "a<custom_token_5>b<custom_token_77>", " <custom_token_9>\n"]
    for i, t in enumerate(texts):
        want = ref.turn_token_into_id(t, i)
        assert sp.parse_custom_token(t, i) == want
        assert tokens.turn_token_into_id(t, i) == want


def test_tokens_decoder_sync_matches(ref):
    strings = sp.synth_token_strings(77, 9)

    async def go():
        async def gen():
            for s in strings:
                yield s
        return [c async for c in ref.tokens_decoder_sync(gen())]
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        want = asyncio.run(go())
    got = list(sp.drop_empty_in_fives(sp.decode_stream(strings, _oracle_convert(ref.model))))
    assert got == want
