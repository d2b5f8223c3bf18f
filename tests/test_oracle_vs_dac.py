"""Independent anchor for the oracle's building blocks.

The third-party `snac` package (the arithmetic of the reference's hot path) is absent, but SNAC's decoder layers
are adapted from the Descript Audio Codec, and the `transformers` wheel in this image ships a DAC
implementation written by other people (`transformers.models.dac.modeling_dac`).  The blocks the two codecs share
- Snake1d, ResidualUnit (Snake -> conv k7 dil d pad 3d -> Snake -> conv k1 -> + x) and DecoderBlock
(Snake -> ConvTranspose1d k=2s stride s pad ceil(s/2) -> RU d=1,3,9) - must agree numerically with the oracle's
restatement once SNAC's additions are neutralised: noise off, and the depthwise k7 convs written into the dense
DAC convs as block-diagonal weights.  CPU only."""
import math

import pytest
import torch

dac = pytest.importorskip("transformers.models.dac.modeling_dac")
from transformers import DacConfig  # noqa: E402

from oracle import snac_ref  # noqa: E402


def _dense_from_depthwise(w):  # [C,1,k] -> [C,C,k] block diagonal
    c, _, k = w.shape
    out = torch.zeros(c, c, k)
    out[torch.arange(c), torch.arange(c)] = w[:, 0]
    return out


def test_snake_matches_dac():
    torch.manual_seed(0)
    mine, theirs = snac_ref.Snake1d(24), dac.Snake1d(24)
    a = 0.5 + torch.rand(1, 24, 1)
    mine.alpha.data.copy_(a)
    theirs.alpha.data.copy_(a)
    x = torch.randn(3, 24, 50) * 3
    assert torch.allclose(mine(x), theirs(x), atol=1e-6)


@pytest.mark.parametrize("stride", [8, 4, 2])
def test_decoder_block_matches_dac(stride):
    torch.manual_seed(stride)
    cin, cout = 32, 16
    cfg = DacConfig(decoder_hidden_size=cin)
    theirs = dac.DacDecoderBlock(cfg, stride=stride, stride_index=0).eval()
    mine = snac_ref.DecoderBlock(cin, cout, stride).eval()
    blk = mine.block
    with torch.no_grad():
        for s in (blk[0], blk[3].block[0], blk[3].block[2], blk[4].block[0], blk[4].block[2], blk[5].block[0], blk[5].block[2]):
            s.alpha.copy_(0.5 + torch.rand_like(s.alpha))
        for conv in [blk[1]] + [ru.block[i] for ru in (blk[3], blk[4], blk[5]) for i in (1, 3)]:
            conv.weight_g.mul_(0.8 + 0.4 * torch.rand_like(conv.weight_g))  # exercise the weight-norm gain
        theirs.snake1.alpha.copy_(blk[0].alpha)
        theirs.conv_t1.weight.copy_(blk[1]._w())     # folded: g * v / ||v|| per INPUT channel
        theirs.conv_t1.bias.copy_(blk[1].bias)
        for ru_m, ru_t in ((blk[3], theirs.res_unit1), (blk[4], theirs.res_unit2), (blk[5], theirs.res_unit3)):
            ru_t.snake1.alpha.copy_(ru_m.block[0].alpha)
            ru_t.conv1.weight.copy_(_dense_from_depthwise(ru_m.block[1]._w()))
            ru_t.conv1.bias.copy_(ru_m.block[1].bias)
            ru_t.snake2.alpha.copy_(ru_m.block[2].alpha)
            ru_t.conv2.weight.copy_(ru_m.block[3]._w())
            ru_t.conv2.bias.copy_(ru_m.block[3].bias)
        blk[2].source = "off"  # NoiseBlock is SNAC's addition
        x = torch.randn(2, cin, 40)
        got, want = mine(x), theirs(x)
    assert got.shape == want.shape == (2, cout, 40 * stride)
    assert theirs.conv_t1.kernel_size == (2 * stride,) and theirs.conv_t1.padding == (math.ceil(stride / 2),)
    assert torch.allclose(got, want, atol=2e-5), float((got - want).abs().max())


@pytest.mark.parametrize("stride", [2, 4, 8])
def test_encoder_block_matches_dac(stride):
    """N4 oracle: EncoderBlock (RU d = 1, 3, 9 -> Snake -> strided conv k = 2s, pad ceil(s/2)) against DacEncoderBlock,
    depthwise convs written into DAC's dense ones as block-diagonal weights."""
    torch.manual_seed(10 + stride)
    cin = 16
    cfg = DacConfig(encoder_hidden_size=cin)
    theirs = dac.DacEncoderBlock(cfg, stride=stride, stride_index=1).eval()   # dimension = 2 * cin: cin -> 2 cin
    mine = snac_ref.EncoderBlock(cin, 2 * cin, stride).eval()
    blk = mine.block
    with torch.no_grad():
        for s in (blk[0].block[0], blk[0].block[2], blk[1].block[0], blk[1].block[2], blk[2].block[0], blk[2].block[2], blk[3]):
            s.alpha.copy_(0.5 + torch.rand_like(s.alpha))
        for conv in [blk[4]] + [ru.block[i] for ru in (blk[0], blk[1], blk[2]) for i in (1, 3)]:
            conv.weight_g.mul_(0.8 + 0.4 * torch.rand_like(conv.weight_g))
        for ru_m, ru_t in ((blk[0], theirs.res_unit1), (blk[1], theirs.res_unit2), (blk[2], theirs.res_unit3)):
            ru_t.snake1.alpha.copy_(ru_m.block[0].alpha)
            ru_t.conv1.weight.copy_(_dense_from_depthwise(ru_m.block[1]._w()))
            ru_t.conv1.bias.copy_(ru_m.block[1].bias)
            ru_t.snake2.alpha.copy_(ru_m.block[2].alpha)
            ru_t.conv2.weight.copy_(ru_m.block[3]._w())
            ru_t.conv2.bias.copy_(ru_m.block[3].bias)
        theirs.snake1.alpha.copy_(blk[3].alpha)
        theirs.conv1.weight.copy_(blk[4]._w())
        theirs.conv1.bias.copy_(blk[4].bias)
        x = torch.randn(2, cin, 16 * stride)
        got, want = mine(x), theirs(x)
    assert got.shape == want.shape
    assert torch.allclose(got, want, atol=2e-5), float((got - want).abs().max())


def test_vector_quantize_matches_dac():
    """N4 oracle: one residual-VQ level (in_proj -> nearest codeword on L2-normalised vectors -> out_proj) against
    DacVectorQuantize: same indices, same quantised latent (stride 1: SNAC's pooling / repeat are its own addition)."""
    torch.manual_seed(3)
    latent, size, dim = 24, 64, 8
    cfg = DacConfig(hidden_size=latent, codebook_size=size, codebook_dim=dim)
    theirs = dac.DacVectorQuantize(cfg).eval()
    mine = snac_ref.VectorQuantize(latent, size, dim, 1).eval()
    with torch.no_grad():
        mine.in_proj.weight_g.mul_(0.8 + 0.4 * torch.rand_like(mine.in_proj.weight_g))
        mine.out_proj.weight_g.mul_(0.8 + 0.4 * torch.rand_like(mine.out_proj.weight_g))
        theirs.in_proj.weight.copy_(mine.in_proj._w()); theirs.in_proj.bias.copy_(mine.in_proj.bias)
        theirs.out_proj.weight.copy_(mine.out_proj._w()); theirs.out_proj.bias.copy_(mine.out_proj.bias)
        theirs.codebook.weight.copy_(mine.codebook.weight)
        z = torch.randn(3, latent, 50)
        z_q, idx = mine.encode(z)
        out = theirs(z)
    # DacVectorQuantize.forward -> (quantized [through out_proj], commitment, codebook loss, indices, projected latents)
    q_t, idx_t = out[0], out[3]
    assert torch.equal(idx, idx_t)
    assert torch.allclose(z_q, q_t, atol=1e-5), float((z_q - q_t).abs().max())
