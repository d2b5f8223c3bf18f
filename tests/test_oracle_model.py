"""Oracle self-checks: the SNAC-24k restatement has the published structure."""
import numpy as np
import torch

from oracle import snac_ref, speechpipe_ref as sp
from project_morpheus_b200 import weights


def test_parameter_counts(oracle_w1):
    dec = sum(p.numel() for p in oracle_w1.decoder.parameters())
    assert dec == 13_012_418
    path = dec + sum(q.codebook.weight.numel() + sum(p.numel() for p in q.out_proj.parameters())
                     for q in oracle_w1.quantizer.quantizers)
    assert path == 13_133_762


def test_output_length_and_range(oracle_w1):
    oracle_w1.set_noise("off")
    for frames in (1, 2, 4):
        codes = [torch.randint(0, 4096, (2, frames * k)) for k in (1, 2, 4)]
        y = oracle_w1.decode(codes)
        assert tuple(y.shape) == (2, 1, 2048 * frames)
        assert float(y.abs().max()) < 1.0


def test_code_4096_raises_like_embedding(oracle_w1):
    codes = [torch.full((1, k), 4096) for k in (1, 2, 4)]
    try:
        oracle_w1.decode(codes)
    except IndexError:
        return
    raise AssertionError("expected IndexError")


def test_noise_injection_is_deterministic_and_matters(oracle_w1):
    codes = [torch.randint(0, 4096, (1, 4 * k), generator=torch.Generator().manual_seed(3)) for k in (1, 2, 4)]
    oracle_w1.set_noise(snac_ref.make_noise(1, 4, seed=1))
    a = oracle_w1.decode(codes)
    oracle_w1.set_noise(snac_ref.make_noise(1, 4, seed=1))
    b = oracle_w1.decode(codes)
    oracle_w1.set_noise(snac_ref.make_noise(1, 4, seed=2))
    c = oracle_w1.decode(codes)
    oracle_w1.set_noise("off")
    assert torch.equal(a, b)
    assert float((a - c).abs().max()) > 1e-3


def test_folding_matches_weight_norm(state_dict_w1, oracle_w1):
    fw = weights.FoldedWeights(state_dict_w1)
    assert fw.num_params() == 13_133_762 - 38 * 0 - sum(
        v.numel() for k, v in weights.normalise_keys(state_dict_w1).items() if k.endswith("weight_g"))
    blk = oracle_w1.decoder.model[2].block
    w_ct = blk[1]._w()  # [Cin, Cout, k], normalised per INPUT channel
    assert torch.allclose(fw.tensors["b0_convt_w"], w_ct, atol=1e-7)
    v = blk[1].weight_v
    manual = blk[1].weight_g * v / v.flatten(1).norm(dim=1).reshape(-1, 1, 1)
    assert torch.allclose(w_ct, manual, atol=1e-6)
    assert torch.allclose(fw.tensors["b0_noise_w"], blk[2].linear._w().reshape(512, 512), atol=1e-7)


def test_both_weight_norm_spellings_load(state_dict_w1):
    renamed = {}
    for k, v in state_dict_w1.items():
        if k.endswith(".weight_g"):
            k = k[:-9] + ".parametrizations.weight.original0"
        elif k.endswith(".weight_v"):
            k = k[:-9] + ".parametrizations.weight.original1"
        renamed[k] = v
    a = weights.FoldedWeights(state_dict_w1).tensors
    b = weights.FoldedWeights(renamed).tensors
    assert all(torch.equal(a[k], b[k]) for k in a)
    snac_ref.SNAC.from_state_dict(renamed)


def test_checkpoint_roundtrip(tmp_path, state_dict_default):
    weights.save_checkpoint(str(tmp_path), state_dict_default)
    back = weights.load_checkpoint(str(tmp_path))
    assert set(back) == set(state_dict_default)
    m = snac_ref.SNAC.from_pretrained(str(tmp_path))
    assert sum(p.numel() for p in m.decoder.parameters()) == 13_012_418


def test_synth_stream_recipe():
    c = sp.synth_codes(0, 4)
    assert c.shape == (28,) and c.min() >= 1 and c.max() <= 4095
    s = sp.synth_token_strings(0, 2)
    assert sp.parse_custom_token(s[8], 8) == int(sp.synth_codes(0, 2)[8])
