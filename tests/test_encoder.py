"""the SNAC encoder.  CPU: the oracle restatement has the published structure (parameter count of
the whole model 19 842 914 = the 19.8 M the model card quotes; output lengths; padding rule).  GPU: the CUDA encode
(through the C ABI) against the oracle - latent to fp32 re-association accuracy, codes equal except at numerical
near-ties, and the decode of the encoded codes within the decode path's own tolerance."""
import numpy as np
import pytest
import torch

from oracle import snac_ref
from project_morpheus_b200 import weights


@pytest.fixture(scope="module")
def full_state_dict(state_dict_w1):
    sd = dict(state_dict_w1)
    sd.update(weights.random_encoder_state_dict(0, "w1"))
    return sd


@pytest.fixture(scope="module")
def oracle_full(full_state_dict):
    torch.set_grad_enabled(False)
    return snac_ref.SNAC.from_state_dict(full_state_dict).eval()


def synth_audio(batch, samples, seed=0):
    g = torch.Generator().manual_seed(seed)
    t = torch.arange(samples) / 24000.0
    x = torch.zeros(batch, 1, samples)
    for b in range(batch):
        f = 110.0 * (b + 1)
        x[b, 0] = 0.4 * torch.sin(2 * np.pi * f * t) + 0.2 * torch.sin(2 * np.pi * 3.1 * f * t + 1.0)
    return x + 0.05 * torch.randn(x.shape, generator=g)


def test_oracle_encoder_structure(oracle_full, state_dict_w1):
    enc = sum(p.numel() for p in oracle_full.encoder.parameters())
    assert enc == 6_690_672  # "6.69 M params"
    total = sum(p.numel() for p in oracle_full.parameters())
    assert total == 19_842_914  # encoder + quantizer + decoder, the published 19.8 M
    # a decode-only dict still loads and has no encoder
    assert snac_ref.SNAC.from_state_dict(state_dict_w1).encoder is None
    x = synth_audio(2, 5000)
    assert oracle_full.preprocess(x).shape[-1] == 6144  # right-padded to a multiple of hop * lcm(4, 1) = 2048
    codes = oracle_full.encode(x)
    assert [tuple(c.shape) for c in codes] == [(2, 3), (2, 6), (2, 12)]
    assert all(int(c.min()) >= 0 and int(c.max()) < 4096 for c in codes)
    z = oracle_full.encode_latent(x)
    assert tuple(z.shape) == (2, 768, 12)


def test_folded_weights_carry_the_encoder(full_state_dict, state_dict_w1):
    fw = weights.FoldedWeights(full_state_dict)
    assert fw.has_encoder and not weights.FoldedWeights(state_dict_w1).has_encoder
    assert tuple(fw.tensors["enc0_down_w"].shape) == (96, 48, 4) and tuple(fw.tensors["enc3_down_w"].shape) == (768, 384, 16)
    assert tuple(fw.tensors["inproj_w2"].shape) == (8, 768)
    # the decode-path tensors are bit-identical with or without the encoder keys (golden vectors depend on them)
    base = weights.FoldedWeights(state_dict_w1)
    assert all(torch.equal(base.tensors[k], fw.tensors[k]) for k in base.tensors)


@pytest.mark.gpu
def test_cuda_encode_matches_oracle(full_state_dict, oracle_full):
    from project_morpheus_b200.engine import SnacEngine
    eng = SnacEngine(full_state_dict, device=0, precision="fp32")
    try:
        for batch, samples in ((3, 24000), (1, 5000), (5, 2048)):
            x = synth_audio(batch, samples, seed=samples)
            want_codes = oracle_full.encode(x)
            want_z = oracle_full.encode_latent(x)
            got_codes, got_z = eng.encode(x, return_latent=True)
            got_z = got_z.cpu()
            assert got_z.shape == want_z.shape
            err = float((got_z - want_z).abs().max())
            assert err <= 2e-4 * max(1.0, float(want_z.abs().max())), err
            total = mism = 0
            for lvl, (g, w) in enumerate(zip(got_codes, want_codes)):
                g = g.cpu()
                assert g.shape == w.shape and g.dtype == torch.int64
                total += w.numel()
                mism += int((g != w).sum())
            # a flipped code is only acceptable at a numerical near-tie; on these inputs there should be (almost) none
            assert mism <= max(1, total // 200), (mism, total)
        # encode -> decode on the GPU equals oracle decode of the same codes (decode tolerance), noise off
        x = synth_audio(2, 8192, seed=5)
        codes = eng.encode(x)
        oracle_full.set_noise("off")
        ref = oracle_full.decode([c.cpu() for c in codes])
        wav = eng.decode_codes(codes, noise="off").cpu()
        assert float((wav - ref).abs().max()) <= 2e-5
    finally:
        eng.close()


@pytest.mark.gpu
def test_snac_shim_encode_and_missing_encoder(full_state_dict, state_dict_w1):
    from project_morpheus_b200 import snac as snac_mod
    m = snac_mod.SNAC.from_state_dict(full_state_dict, precision="fp32", noise="off").to("cuda")
    codes = m.encode(synth_audio(1, 4096))
    assert [tuple(c.shape) for c in codes] == [(1, 2), (1, 4), (1, 8)]
    y = m.decode(codes)
    assert tuple(y.shape) == (1, 1, 4096)
    m2 = snac_mod.SNAC.from_state_dict(state_dict_w1, precision="fp32", noise="off").to("cuda")
    with pytest.raises(Exception, match="encoder"):
        m2.encode(synth_audio(1, 4096))
