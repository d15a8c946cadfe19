"""Host logic of the VAE drop-in (reptext_b200/vae.py) WITHOUT a GPU: the operator entry points are replaced by the torch
statement of their contracts (tests/ops_emulator.py) and the result is compared with the oracle in fp32.  This pins the
weight packing (tap-major 3x3 kernels, padded channels), the im2col / Downsample2D padding convention, the folded
attention bias, the transposed-V trick and the block order; the kernels themselves are checked on the GPU
(tests/test_vae_gpu.py)."""
import pytest
import torch

import ops_emulator
from oracle import vae_oracle as V
from reptext_b200 import ops, vae
from reptext_b200.models import FrozenConfig
from util import rel_l2


def _cpu_model(cfg, sd, conv_impl):
    m = object.__new__(vae.AutoencoderKL)      # the constructor insists on a CUDA device; the emulator runs on the CPU
    full = dict(vae.FLUX_VAE_CONFIG)
    full.update(cfg)
    m.config = FrozenConfig(**full)
    m.dtype, m.device, m.conv_impl, m._w = torch.float32, torch.device("cpu"), conv_impl, {}
    m._prepare(sd)
    return m


@pytest.mark.parametrize("conv_impl", ["implicit", "im2col"])
def test_vae_host_logic_matches_oracle(monkeypatch, conv_impl):
    ops_emulator.install(monkeypatch, ops)
    cfg = dict(V.FLUX_VAE_CONFIG, block_out_channels=(64, 128, 128, 128))
    sd = V.random_state_dict(cfg, seed=3)
    m = _cpu_model(cfg, sd, conv_impl)
    g = torch.Generator().manual_seed(4)
    img = torch.rand(2, 3, 128, 128, generator=g) * 2 - 1
    with torch.no_grad():
        ref_mom = V.encode_moments(sd, cfg, img)
    post = m.encode(img).latent_dist
    assert rel_l2(post.parameters, ref_mom) < 1e-4
    noise = torch.randn(2, 16, 16, 16, generator=g)
    assert rel_l2(post.sample_with_noise(noise), V.sample_posterior(ref_mom, noise)) < 1e-4
    assert rel_l2(post.mode(), ref_mom[:, :16]) < 1e-4
    z = torch.randn(2, 16, 16, 16, generator=g)
    with torch.no_grad():
        ref_img = V.decode(sd, cfg, z)
    out = m.decode(z, return_dict=False)[0]
    assert out.shape == (2, 3, 128, 128)
    assert rel_l2(out, ref_img) < 1e-4


def test_vae_rejects_cpu_and_bad_shapes():
    cfg = dict(V.FLUX_VAE_CONFIG, block_out_channels=(64, 128, 128, 128))
    with pytest.raises(ValueError):
        vae.AutoencoderKL(cfg, {}, device="cpu")                       # no CPU path
    with pytest.raises(ValueError):
        vae.AutoencoderKL(dict(cfg, block_out_channels=(96, 128)), {}, device="cuda")
    with pytest.raises(ValueError):
        vae.AutoencoderKL._check_hw((24, 24))
    vae.AutoencoderKL._check_hw((128, 128))
    vae.AutoencoderKL._check_hw((8, 16))
    with pytest.raises(ValueError):
        vae.AutoencoderKL._check_hw((4, 16))


def test_pack_conv3x3_weight_layout():
    w = torch.arange(2 * 3 * 9, dtype=torch.float32).view(2, 3, 3, 3)
    p = ops.pack_conv3x3_weight(w)
    assert p.shape == (64, 9 * 64)
    for ky in range(3):
        for kx in range(3):
            assert torch.equal(p[:2, (ky * 3 + kx) * 64:(ky * 3 + kx) * 64 + 3], w[:, :, ky, kx])
    assert float(p[2:].abs().sum()) == 0 and float(p[:, 3:64].abs().sum()) == 0
