"""CPU: structural checks of the VAE oracle (oracle/vae_oracle.py — third-party architecture, "parity unpinned": the reference
holds no golden for diffusers' AutoencoderKL). What can be pinned without the package: the layer inventory / key names / shapes of
the published FLUX.1 VAE checkpoint layout, the `Downsample2D` padding convention, the single-head attention against torch's
SDPA, and the sampling formula."""
import math

import torch
import torch.nn.functional as F

from oracle import vae_oracle as O


def test_state_dict_inventory_matches_the_published_layout():
    sd = O.init_state_dict(O.VAEConfig.flux())
    # 4 levels x 2 resnets (encoder) / x 3 resnets (decoder), 3 down- / up-samplers, two mid blocks with attention
    assert sum(k.endswith("conv1.weight") for k in sd) == 4 * 2 + 4 * 3 + 2 * 2
    assert sum("downsamplers" in k for k in sd) == 3 * 2 and sum("upsamplers" in k for k in sd) == 3 * 2
    assert sum("conv_shortcut.weight" in k for k in sd) == 2 + 2  # 128->256, 256->512 down; 512->256, 256->128 up
    assert sd["encoder.conv_in.weight"].shape == (128, 3, 3, 3) and sd["encoder.conv_out.weight"].shape == (32, 512, 3, 3)
    assert sd["decoder.conv_in.weight"].shape == (512, 16, 3, 3) and sd["decoder.conv_out.weight"].shape == (3, 128, 3, 3)
    assert sd["decoder.up_blocks.2.resnets.0.conv_shortcut.weight"].shape == (256, 512, 1, 1)
    assert sd["encoder.mid_block.attentions.0.to_q.weight"].shape == (512, 512)
    assert sd["decoder.mid_block.attentions.0.group_norm.weight"].shape == (512,)
    assert not any(k.startswith(("quant_conv", "post_quant_conv")) for k in sd)
    n_params = sum(v.numel() for v in sd.values())
    assert 83_000_000 < n_params < 84_500_000  # the published FLUX.1 / SD3 16-channel VAE: 83.8 M parameters


def test_shapes_and_determinism_tiny():
    cfg = O.VAEConfig.tiny()
    orc = O.VAEOracle(cfg, O.init_state_dict(cfg, seed=1))
    g = torch.Generator().manual_seed(0)
    img = torch.rand(2, 3, 32, 48, generator=g) * 2 - 1
    noise = torch.randn(2, 16, 16, 24, generator=g)
    lat = orc.encode(img, noise)
    assert lat.shape == (2, 16, 16, 24)
    assert torch.equal(lat, orc.encode(img, noise))
    assert orc.decode(lat).shape == (2, 3, 32, 48)
    mean, logvar = orc.encoder(img).chunk(2, dim=1)
    want = (mean + torch.exp(0.5 * logvar.clamp(-30, 20)) * noise - cfg.shift_factor) * cfg.scaling_factor
    torch.testing.assert_close(lat, want)
    torch.testing.assert_close(orc.encode(img, None), (mean - cfg.shift_factor) * cfg.scaling_factor)


def test_downsample_pads_bottom_and_right_only():
    cfg = O.VAEConfig.tiny()
    sd = O.init_state_dict(cfg, seed=2)
    orc = O.VAEOracle(cfg, sd)
    orc.record = True
    img = torch.rand(1, 3, 8, 8) * 2 - 1
    orc.encoder(img)
    # recompute level 0 by hand: the stride-2 convolution sees the un-padded top-left corner first
    h = F.conv2d(img, sd["encoder.conv_in.weight"], sd["encoder.conv_in.bias"], padding=1)
    h = O.resnet_block(sd, "encoder.down_blocks.0.resnets.0", h, cfg.norm_num_groups)
    w, b = sd["encoder.down_blocks.0.downsamplers.0.conv.weight"], sd["encoder.down_blocks.0.downsamplers.0.conv.bias"]
    corner = (h[0, :, 0:3, 0:3] * w).sum(dim=(1, 2, 3)) + b
    torch.testing.assert_close(orc.trace["encoder.down_blocks.0"][0, :, 0, 0], corner, rtol=1e-4, atol=1e-5)
    assert orc.trace["encoder.down_blocks.0"].shape == (1, 64, 4, 4)


def test_attention_block_is_one_head_sdpa_with_residual():
    c, g = 64, 32
    gen = torch.Generator().manual_seed(3)
    sd = {}
    O._gn_init(sd, "a.group_norm", c, gen)
    for n in ("to_q", "to_k", "to_v", "to_out.0"):
        O._lin_init(sd, f"a.{n}", c, c, gen)
    x = torch.randn(2, c, 5, 6, generator=gen)
    t = F.group_norm(x, g, sd["a.group_norm.weight"], sd["a.group_norm.bias"], eps=1e-6).flatten(2).transpose(1, 2)
    q, k, v = (F.linear(t, sd[f"a.{n}.weight"], sd[f"a.{n}.bias"]) for n in ("to_q", "to_k", "to_v"))
    o = F.scaled_dot_product_attention(q[:, None], k[:, None], v[:, None], scale=1 / math.sqrt(c))[:, 0]
    want = x + F.linear(o, sd["a.to_out.0.weight"], sd["a.to_out.0.bias"]).transpose(1, 2).reshape(2, c, 5, 6)
    torch.testing.assert_close(O.attention_block(sd, "a", x, g), want, rtol=1e-4, atol=1e-5)
