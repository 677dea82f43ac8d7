"""ORACLE — test infrastructure, NOT product code.

CPU-runnable, pure-torch restatement of the VAE the reference pipelines run around the denoiser (SURVEY.md §8 (f)4 tail):
`vae.encode(control_image).latent_dist.sample()` followed by `(x - shift_factor) * scaling_factor`
(src/UniGenPipeline.py:306-308, :635-636, :960-961; src/condition.py:97-100) and
`vae.decode(latents / scaling_factor + shift_factor)` (:439-441, :797-798, :1124-1125).
Only tests/, `__graft_entry__.smoke()` and `bench.py`'s baseline legs may import this file.

Parity pinning status: "parity unpinned". The network is third-party — diffusers 0.32.2 `AutoencoderKL`
(`models/autoencoders/autoencoder_kl.py`, `vae.py::{Encoder, Decoder, DiagonalGaussianDistribution}`,
`unets/unet_2d_blocks.py::{DownEncoderBlock2D, UpDecoderBlock2D, UNetMidBlock2D}`, `resnet.py::ResnetBlock2D`,
`downsampling.py::Downsample2D`, `upsampling.py::Upsample2D`, `attention_processor.py::Attention` as the deprecated
attention block) — absent from /root/reference, this image and the wheelhouse; the reference holds no test or golden for
it. Restated from the published architecture:

  Encoder   conv_in 3x3 -> per level [ResnetBlock2D x layers_per_block, Downsample2D (pad (0,1,0,1), conv 3x3 stride 2) except
            the last] -> mid (resnet, attention, resnet) -> GroupNorm(32, eps 1e-6) -> SiLU -> conv_out 3x3 to 2 * latent_channels
  Decoder   conv_in 3x3 -> mid (resnet, attention, resnet) -> per level, channels reversed [ResnetBlock2D x (layers_per_block + 1),
            Upsample2D (nearest x2, conv 3x3) except the last] -> GroupNorm -> SiLU -> conv_out 3x3
  ResnetBlock2D (temb None, output_scale_factor 1): x + conv2(silu(gn2(conv1(silu(gn1(x)))))), 1x1 `conv_shortcut` on x when the
            channel count changes; GroupNorm eps 1e-6
  Attention (mid block): residual + to_out(softmax(q k^T / sqrt(C)) v) with ONE head of width C, q / k / v = Linear(gn(x)),
            tokens = pixels; rescale_output_factor 1
  DiagonalGaussianDistribution: mean, logvar = chunk(moments, 2, dim=1); logvar clamped to [-30, 20]; sample = mean +
            exp(0.5 * logvar) * randn; mode = mean
FLUX.1 / SD3.5 VAE config: latent_channels 16, block_out_channels (128, 256, 512, 512), layers_per_block 2, no quant / post-quant
conv; scaling / shift factors 0.3611 / 0.1159 (Flux) and 1.5305 / 0.0609 (SD3).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


@dataclass
class VAEConfig:
    in_channels: int = 3
    out_channels: int = 3
    latent_channels: int = 16
    block_out_channels: Tuple[int, ...] = (128, 256, 512, 512)
    layers_per_block: int = 2
    norm_num_groups: int = 32
    scaling_factor: float = 0.3611
    shift_factor: float = 0.1159
    mid_block_add_attention: bool = True

    @staticmethod
    def tiny() -> "VAEConfig":
        """Two levels (one down / up-sampling), 64 -> 128 channels: every block type of the full model at test size."""
        return VAEConfig(block_out_channels=(64, 128), layers_per_block=1, norm_num_groups=32)

    @staticmethod
    def flux() -> "VAEConfig":
        return VAEConfig()

    @staticmethod
    def sd3() -> "VAEConfig":
        return VAEConfig(scaling_factor=1.5305, shift_factor=0.0609)


def _conv(sd, p: str, x: Tensor, stride: int = 1, padding: int = 1) -> Tensor:
    return F.conv2d(x, sd[p + ".weight"], sd[p + ".bias"], stride=stride, padding=padding)


def _gn(sd, p: str, x: Tensor, groups: int) -> Tensor:
    return F.group_norm(x, groups, sd[p + ".weight"], sd[p + ".bias"], eps=1e-6)


def resnet_block(sd, p: str, x: Tensor, groups: int) -> Tensor:
    h = _conv(sd, p + ".conv1", F.silu(_gn(sd, p + ".norm1", x, groups)))
    h = _conv(sd, p + ".conv2", F.silu(_gn(sd, p + ".norm2", h, groups)))
    if p + ".conv_shortcut.weight" in sd:
        x = _conv(sd, p + ".conv_shortcut", x, padding=0)
    return x + h


def attention_block(sd, p: str, x: Tensor, groups: int) -> Tensor:
    B, Cc, H, W = x.shape
    t = _gn(sd, p + ".group_norm", x.view(B, Cc, H * W), groups).transpose(1, 2)  # [B, HW, C]
    q = F.linear(t, sd[p + ".to_q.weight"], sd[p + ".to_q.bias"])
    k = F.linear(t, sd[p + ".to_k.weight"], sd[p + ".to_k.bias"])
    v = F.linear(t, sd[p + ".to_v.weight"], sd[p + ".to_v.bias"])
    a = torch.softmax(q @ k.transpose(1, 2) / math.sqrt(Cc), dim=-1) @ v  # one head of width C
    o = F.linear(a, sd[p + ".to_out.0.weight"], sd[p + ".to_out.0.bias"])
    return x + o.transpose(1, 2).reshape(B, Cc, H, W)


def mid_block(sd, p: str, x: Tensor, cfg: VAEConfig) -> Tensor:
    x = resnet_block(sd, p + ".resnets.0", x, cfg.norm_num_groups)
    if cfg.mid_block_add_attention:
        x = attention_block(sd, p + ".attentions.0", x, cfg.norm_num_groups)
    return resnet_block(sd, p + ".resnets.1", x, cfg.norm_num_groups)


class VAEOracle:
    def __init__(self, cfg: VAEConfig, state_dict: Dict[str, Tensor]):
        self.cfg, self.sd = cfg, state_dict
        self.trace: Dict[str, Tensor] = {}
        self.record = False

    def _rec(self, name: str, t: Tensor) -> None:
        if self.record:
            self.trace[name] = t.detach().clone()

    def encoder_stages(self):
        """The encoder as an ordered list of (trace name, stage function): stage i maps the output of stage i - 1 (the image for
        stage 0) to the tensor recorded under its name — lets a test feed every stage the PRODUCT's own input of that stage."""
        cfg, sd, g = self.cfg, self.sd, self.cfg.norm_num_groups
        n = len(cfg.block_out_channels)
        stages = [("encoder.conv_in", lambda x: _conv(sd, "encoder.conv_in", x))]

        def down(i):
            def run(h):
                for j in range(cfg.layers_per_block):
                    h = resnet_block(sd, f"encoder.down_blocks.{i}.resnets.{j}", h, g)
                if i < n - 1:  # Downsample2D(padding=0): F.pad(x, (0, 1, 0, 1)) then the stride-2 convolution
                    h = _conv(sd, f"encoder.down_blocks.{i}.downsamplers.0.conv", F.pad(h, (0, 1, 0, 1)), stride=2, padding=0)
                return h
            return run

        stages += [(f"encoder.down_blocks.{i}", down(i)) for i in range(n)]
        stages.append(("encoder.mid_block", lambda h: mid_block(sd, "encoder.mid_block", h, cfg)))
        stages.append(("encoder.moments", lambda h: _conv(sd, "encoder.conv_out", F.silu(_gn(sd, "encoder.conv_norm_out", h, g)))))
        return stages

    def decoder_stages(self):
        cfg, sd, g = self.cfg, self.sd, self.cfg.norm_num_groups
        n = len(cfg.block_out_channels)
        stages = [("decoder.conv_in", lambda z: _conv(sd, "decoder.conv_in", z)),
                  ("decoder.mid_block", lambda h: mid_block(sd, "decoder.mid_block", h, cfg))]

        def up(i):
            def run(h):
                for j in range(cfg.layers_per_block + 1):
                    h = resnet_block(sd, f"decoder.up_blocks.{i}.resnets.{j}", h, g)
                if i < n - 1:  # Upsample2D: nearest x2, then the convolution
                    h = _conv(sd, f"decoder.up_blocks.{i}.upsamplers.0.conv", F.interpolate(h, scale_factor=2.0, mode="nearest"))
                return h
            return run

        stages += [(f"decoder.up_blocks.{i}", up(i)) for i in range(n)]
        stages.append(("decoder.sample", lambda h: _conv(sd, "decoder.conv_out", F.silu(_gn(sd, "decoder.conv_norm_out", h, g)))))
        return stages

    def encoder(self, x: Tensor) -> Tensor:
        for name, fn in self.encoder_stages():
            x = fn(x)
            self._rec(name, x)
        return x

    def decoder(self, z: Tensor) -> Tensor:
        for name, fn in self.decoder_stages():
            z = fn(z)
            self._rec(name, z)
        return z

    def encode(self, x: Tensor, noise: Optional[Tensor] = None) -> Tensor:
        """`vae.encode(x).latent_dist.sample()` (noise given; `.mode()` for None), then `(z - shift) * scaling` — the
        latents as the pipelines hand them to the transformer."""
        mean, logvar = self.encoder(x).chunk(2, dim=1)
        z = mean
        if noise is not None:
            z = mean + torch.exp(0.5 * logvar.clamp(-30.0, 20.0)) * noise
        return (z - self.cfg.shift_factor) * self.cfg.scaling_factor

    def decode(self, latents: Tensor) -> Tensor:
        """`vae.decode(latents / scaling + shift)` on transformer-side latents."""
        return self.decoder(latents / self.cfg.scaling_factor + self.cfg.shift_factor)


# ------------------------------------------------------------------------------------------------------------------
def _conv_init(sd, p, c_out, c_in, k, gen):
    bound = 1.0 / math.sqrt(c_in * k * k)
    sd[p + ".weight"] = (torch.rand(c_out, c_in, k, k, generator=gen) * 2 - 1) * bound
    sd[p + ".bias"] = (torch.rand(c_out, generator=gen) * 2 - 1) * bound


def _lin_init(sd, p, c_out, c_in, gen):
    bound = 1.0 / math.sqrt(c_in)
    sd[p + ".weight"] = (torch.rand(c_out, c_in, generator=gen) * 2 - 1) * bound
    sd[p + ".bias"] = (torch.rand(c_out, generator=gen) * 2 - 1) * bound


def _gn_init(sd, p, c, gen):
    sd[p + ".weight"] = 1.0 + 0.1 * torch.randn(c, generator=gen)
    sd[p + ".bias"] = 0.1 * torch.randn(c, generator=gen)


def _resnet_init(sd, p, c_in, c_out, gen):
    _gn_init(sd, p + ".norm1", c_in, gen)
    _conv_init(sd, p + ".conv1", c_out, c_in, 3, gen)
    _gn_init(sd, p + ".norm2", c_out, gen)
    _conv_init(sd, p + ".conv2", c_out, c_out, 3, gen)
    if c_in != c_out:
        _conv_init(sd, p + ".conv_shortcut", c_out, c_in, 1, gen)


def _mid_init(sd, p, c, cfg: VAEConfig, gen):
    _resnet_init(sd, p + ".resnets.0", c, c, gen)
    if cfg.mid_block_add_attention:
        _gn_init(sd, p + ".attentions.0.group_norm", c, gen)
        for n in ("to_q", "to_k", "to_v", "to_out.0"):
            _lin_init(sd, f"{p}.attentions.0.{n}", c, c, gen)
    _resnet_init(sd, p + ".resnets.1", c, c, gen)


def init_state_dict(cfg: VAEConfig, seed: int = 0) -> Dict[str, Tensor]:
    """Deterministic random weights under the diffusers AutoencoderKL state-dict names."""
    gen = torch.Generator().manual_seed(seed)
    sd: Dict[str, Tensor] = {}
    ch = cfg.block_out_channels
    _conv_init(sd, "encoder.conv_in", ch[0], cfg.in_channels, 3, gen)
    c_prev = ch[0]
    for i, c in enumerate(ch):
        for j in range(cfg.layers_per_block):
            _resnet_init(sd, f"encoder.down_blocks.{i}.resnets.{j}", c_prev if j == 0 else c, c, gen)
        if i < len(ch) - 1:
            _conv_init(sd, f"encoder.down_blocks.{i}.downsamplers.0.conv", c, c, 3, gen)
        c_prev = c
    _mid_init(sd, "encoder.mid_block", ch[-1], cfg, gen)
    _gn_init(sd, "encoder.conv_norm_out", ch[-1], gen)
    _conv_init(sd, "encoder.conv_out", 2 * cfg.latent_channels, ch[-1], 3, gen)
    rev = tuple(reversed(ch))
    _conv_init(sd, "decoder.conv_in", rev[0], cfg.latent_channels, 3, gen)
    _mid_init(sd, "decoder.mid_block", rev[0], cfg, gen)
    c_prev = rev[0]
    for i, c in enumerate(rev):
        for j in range(cfg.layers_per_block + 1):
            _resnet_init(sd, f"decoder.up_blocks.{i}.resnets.{j}", c_prev if j == 0 else c, c, gen)
        if i < len(rev) - 1:
            _conv_init(sd, f"decoder.up_blocks.{i}.upsamplers.0.conv", c, c, 3, gen)
        c_prev = c
    _gn_init(sd, "decoder.conv_norm_out", rev[-1], gen)
    _conv_init(sd, "decoder.conv_out", cfg.out_channels, rev[-1], 3, gen)
    return sd
