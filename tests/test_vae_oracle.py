"""The VAE oracle (oracle/vae_oracle.py, diffusers AutoencoderKL restated) pinned against an INDEPENDENT implementation on
this box: the Black-Forest-Labs autoencoder shipped with torchtitan.  A BFL module with its own random initialisation is
exported into diffusers' parameter names and both sides run on the same input."""
import pytest
import torch

from oracle import vae_oracle as V
from util import rel_l2

ae = pytest.importorskip("torchtitan.experiments.flux.model.autoencoder")

CFG = dict(V.FLUX_VAE_CONFIG, block_out_channels=(32, 64, 128, 128))


def _bfl():
    torch.manual_seed(0)
    m = ae.AutoEncoder(ae.AutoEncoderParams(resolution=64, in_channels=3, ch=32, out_ch=3, ch_mult=(1, 2, 4, 4),
                                            num_res_blocks=2, z_channels=16))
    with torch.no_grad():      # default inits leave the norms at (1, 0): perturb them so the affine part is exercised
        for mod in m.modules():
            if isinstance(mod, torch.nn.GroupNorm):
                mod.weight.add_(0.1 * torch.randn_like(mod.weight))
                mod.bias.add_(0.1 * torch.randn_like(mod.bias))
    return m.eval()


def _to_diffusers_names(m) -> dict:
    sd = {}

    def put(name, mod):
        sd[name + ".weight"] = mod.weight.detach().clone()
        sd[name + ".bias"] = mod.bias.detach().clone()

    def resnet(name, blk):
        put(name + "norm1", blk.norm1); put(name + "conv1", blk.conv1)
        put(name + "norm2", blk.norm2); put(name + "conv2", blk.conv2)
        if blk.in_channels != blk.out_channels:
            put(name + "conv_shortcut", blk.nin_shortcut)

    def mid(name, mm):
        resnet(name + "mid_block.resnets.0.", mm.block_1)
        resnet(name + "mid_block.resnets.1.", mm.block_2)
        a = name + "mid_block.attentions.0."
        put(a + "group_norm", mm.attn_1.norm)
        for dn, conv in (("to_q", mm.attn_1.q), ("to_k", mm.attn_1.k), ("to_v", mm.attn_1.v), ("to_out.0", mm.attn_1.proj_out)):
            sd[a + dn + ".weight"] = conv.weight.detach()[:, :, 0, 0].clone()     # 1x1 convolution -> linear
            sd[a + dn + ".bias"] = conv.bias.detach().clone()

    e, d = m.encoder, m.decoder
    put("encoder.conv_in", e.conv_in)
    for i, lvl in enumerate(e.down):
        for j, blk in enumerate(lvl.block):
            resnet(f"encoder.down_blocks.{i}.resnets.{j}.", blk)
        if hasattr(lvl, "downsample"):
            put(f"encoder.down_blocks.{i}.downsamplers.0.conv", lvl.downsample.conv)
    mid("encoder.", e.mid)
    put("encoder.conv_norm_out", e.norm_out); put("encoder.conv_out", e.conv_out)
    put("decoder.conv_in", d.conv_in)
    mid("decoder.", d.mid)
    n = len(d.up)
    for i in range(n):                       # diffusers' up_blocks.0 runs first = BFL's up[n - 1]
        lvl = d.up[n - 1 - i]
        for j, blk in enumerate(lvl.block):
            resnet(f"decoder.up_blocks.{i}.resnets.{j}.", blk)
        if hasattr(lvl, "upsample"):
            put(f"decoder.up_blocks.{i}.upsamplers.0.conv", lvl.upsample.conv)
    put("decoder.conv_norm_out", d.norm_out); put("decoder.conv_out", d.conv_out)
    return sd


def test_parameter_table_matches_the_exported_names():
    sd = _to_diffusers_names(_bfl())
    want = V.param_shapes(CFG)
    assert set(sd) == set(want)
    assert all(tuple(sd[k].shape) == tuple(want[k]) for k in want)


def test_encoder_and_decoder_match_bfl():
    m = _bfl()
    sd = _to_diffusers_names(m)
    g = torch.Generator().manual_seed(1)
    x = torch.rand(2, 3, 64, 48, generator=g) * 2 - 1
    z = torch.randn(2, 16, 8, 6, generator=g)
    with torch.no_grad():
        assert rel_l2(V.encode_moments(sd, CFG, x), m.encoder(x)) < 2e-5
        assert rel_l2(V.decode(sd, CFG, z), m.decoder(z)) < 2e-5
        # the scale / shift conventions of the pipelines (pipeline_flux_controlnet.py:705-708, :1137-1139)
        noise = torch.randn(2, 16, 8, 6, generator=g)
        mean, logvar = m.encoder(x).chunk(2, dim=1)
        ref = CFG["scaling_factor"] * (mean + torch.exp(0.5 * logvar) * noise - CFG["shift_factor"])
        assert rel_l2(V.encode_for_pipeline(sd, CFG, x, noise), ref) < 2e-5
        assert rel_l2(V.decode_for_pipeline(sd, CFG, z), m.decoder(z / CFG["scaling_factor"] + CFG["shift_factor"])) < 2e-5


def test_random_state_dict_is_well_scaled():
    sd = V.random_state_dict(CFG, seed=3)
    with torch.no_grad():
        y = V.decode(sd, CFG, torch.randn(1, 16, 8, 8, generator=torch.Generator().manual_seed(2)))
    assert y.shape == (1, 3, 64, 64) and torch.isfinite(y).all() and 1e-3 < float(y.std()) < 1e3
