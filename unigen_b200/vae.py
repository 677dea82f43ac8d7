"""Host-side mirror of the VAE the reference pipelines run around the denoiser (diffusers 0.32.2 `AutoencoderKL`) for the
B200-native path — SURVEY.md §8 (f)4 tail.

Reference call sites: `vae.encode(control_image).latent_dist.sample()` then `(x - shift_factor) * scaling_factor`
(src/UniGenPipeline.py:306-308, :635-636, :960-961; src/condition.py:97-100) and
`vae.decode(latents / scaling_factor + shift_factor, return_dict=False)[0]` (:439-441, :797-798, :1124-1125).

Public surface kept: `AutoencoderKL(**config)`, `.config.{scaling_factor, shift_factor, latent_channels, block_out_channels}`,
`.encode(x).latent_dist.{sample(generator), mode()}`, `.decode(z, return_dict)`, `.dtype`, diffusers state-dict keys
(`encoder.down_blocks.0.resnets.0.conv1.weight`, ... with nn.Conv2d's [co, ci, kh, kw] shapes: views into the GEMM layout).
Every arithmetic op runs in libunigen_b200.so; there is no eager / CPU fallback.

Data layout in HBM: activations NHWC bf16 (a pixel is one GEMM row), convolution weights as [co, (ky, kx, ci)] matrices.
  * 3x3 / stride 1 convolutions = implicit GEMM (`ug_conv3x3_bf16`: the TMA unit gathers the A tile per filter tap, zero-fills the
    halo; bias / residual add in the GEMM epilogue) — every resnet convolution and the up-sampler convolutions;
  * stride-2 `Downsample2D`, the 3-channel / 16-channel stems and widths the implicit path does not tile = patch gather
    (`ug_im2col_bf16`) + the plain GEMM; 1x1 `conv_shortcut` and the attention projections = the plain GEMM over pixel rows;
  * GroupNorm(+SiLU), nearest x2 up-sampling, the 512-wide single-head attention's row softmax: HBM-bound kernels (ug_vae.cu);
    the attention itself = GEMM (scaled q k^T, bf16 scores) -> softmax -> GEMM (p v^T'), v produced already transposed.
"""
from __future__ import annotations

import math
import types
from typing import Any, Optional, Sequence

import torch

from . import ops
from .model import BF16, _DenoiserBase, _Weights


class _ConvW:
    """nn.Conv2d(c_in, c_out, k) parameters as the GEMM matrix [c_out_pad, k_pad] (column (ky * k + kx) * c_in + ci); the state-dict
    entry is the [c_out, c_in, k, k] VIEW of it. Padding rows / columns are zeros and stay zeros."""

    def __init__(self, ws: _Weights, name: str, c_out: int, c_in: int, k: int):
        self.c_in, self.c_out, self.k = c_in, c_out, k
        self.k_cols = k * k * c_in
        # both GEMM dimensions cover at least one 64-element TMA box row (the 3- / 16-channel stems, the 3- / 32-channel heads)
        self.k_pad = (self.k_cols + 63) // 64 * 64
        self.c_out_pad = max(64, (c_out + 7) // 8 * 8)
        self.w = ws.alloc(self.c_out_pad, self.k_pad)
        self.b = ws.alloc(self.c_out_pad)
        ws.views[name + ".weight"] = self.w[:c_out, :self.k_cols].view(c_out, k, k, c_in).permute(0, 3, 1, 2)
        ws.views[name + ".bias"] = self.b[:c_out]


class _NormW:
    def __init__(self, ws: _Weights, name: str, c: int):
        self.w, self.b = ws.alloc(c), ws.alloc(c)
        ws.views[name + ".weight"], ws.views[name + ".bias"] = self.w, self.b


class _ResnetW:
    def __init__(self, ws: _Weights, p: str, c_in: int, c_out: int):
        self.norm1, self.conv1 = _NormW(ws, p + ".norm1", c_in), _ConvW(ws, p + ".conv1", c_out, c_in, 3)
        self.norm2, self.conv2 = _NormW(ws, p + ".norm2", c_out), _ConvW(ws, p + ".conv2", c_out, c_out, 3)
        self.shortcut = _ConvW(ws, p + ".conv_shortcut", c_out, c_in, 1) if c_in != c_out else None


class _AttnW:
    def __init__(self, ws: _Weights, p: str, c: int):
        self.norm = _NormW(ws, p + ".group_norm", c)
        self.q, self.k, self.v = (ws.linear(f"{p}.{n}", c, c) for n in ("to_q", "to_k", "to_v"))
        self.out = ws.linear(p + ".to_out.0", c, c)


class _MidW:
    def __init__(self, ws: _Weights, p: str, c: int, attention: bool):
        self.res0 = _ResnetW(ws, p + ".resnets.0", c, c)
        self.attn = _AttnW(ws, p + ".attentions.0", c) if attention else None
        self.res1 = _ResnetW(ws, p + ".resnets.1", c, c)


class DiagonalGaussianDistribution:
    """diffusers `DiagonalGaussianDistribution` over the encoder's moments (kept on the device as NHWC bf16 [B, h, w, 2c])."""

    def __init__(self, moments_nhwc: torch.Tensor, latent_channels: int):
        self._m, self._c = moments_nhwc, latent_channels

    def sample(self, generator: Optional[torch.Generator] = None, shift: float = 0.0, scale: float = 1.0) -> torch.Tensor:
        """mean + exp(0.5 * clamp(logvar, -30, 20)) * randn -> NCHW bf16 [B, c, h, w]; `shift` / `scale` fuse the pipelines'
        `(z - shift_factor) * scaling_factor` into the same pass."""
        B, h, w, _ = self._m.shape
        gdev = generator.device if generator is not None else self._m.device
        noise = torch.randn((B, self._c, h, w), generator=generator, device=gdev, dtype=torch.float32).to(self._m.device)
        return ops.vae_sample(self._m, self._c, noise, shift, scale)

    def mode(self, shift: float = 0.0, scale: float = 1.0) -> torch.Tensor:
        return ops.vae_sample(self._m, self._c, None, shift, scale)

    @property
    def mean(self) -> torch.Tensor:
        return self.mode()


class AutoencoderKL(_DenoiserBase):
    """B200-native drop-in for diffusers `AutoencoderKL` as the reference pipelines use it (encode the condition image, decode
    the final latents)."""

    def __init__(self, in_channels: int = 3, out_channels: int = 3, latent_channels: int = 16,
                 block_out_channels: Sequence[int] = (128, 256, 512, 512), layers_per_block: int = 2, norm_num_groups: int = 32,
                 scaling_factor: float = 0.3611, shift_factor: float = 0.1159, use_quant_conv: bool = False,
                 use_post_quant_conv: bool = False, mid_block_add_attention: bool = True, device: Any = "cuda", **unused):
        super().__init__()
        self._init_base(device)
        if use_quant_conv or use_post_quant_conv:
            raise ops.UgError("AutoencoderKL (B200-native): quant_conv / post_quant_conv are not built (the FLUX.1 and SD3.5 VAEs "
                              "ship with use_quant_conv = use_post_quant_conv = False)")
        ch = tuple(int(c) for c in block_out_channels)
        if any(c % 64 or c % norm_num_groups for c in ch):
            raise ops.UgError("block_out_channels must be multiples of 64 (one TMA box row of channels) and of norm_num_groups")
        self.config = types.SimpleNamespace(
            in_channels=in_channels, out_channels=out_channels, latent_channels=latent_channels, block_out_channels=ch,
            layers_per_block=layers_per_block, norm_num_groups=norm_num_groups, scaling_factor=scaling_factor, shift_factor=shift_factor,
            use_quant_conv=False, use_post_quant_conv=False, mid_block_add_attention=mid_block_add_attention)
        ws, n = self._ws, len(ch)
        # ---- Encoder ----
        self.enc_in = _ConvW(ws, "encoder.conv_in", ch[0], in_channels, 3)
        self.enc_down = []
        c_prev = ch[0]
        for i, c in enumerate(ch):
            res = [_ResnetW(ws, f"encoder.down_blocks.{i}.resnets.{j}", c_prev if j == 0 else c, c) for j in range(layers_per_block)]
            down = _ConvW(ws, f"encoder.down_blocks.{i}.downsamplers.0.conv", c, c, 3) if i < n - 1 else None
            self.enc_down.append((res, down))
            c_prev = c
        self.enc_mid = _MidW(ws, "encoder.mid_block", ch[-1], mid_block_add_attention)
        self.enc_norm_out = _NormW(ws, "encoder.conv_norm_out", ch[-1])
        self.enc_out = _ConvW(ws, "encoder.conv_out", 2 * latent_channels, ch[-1], 3)
        # ---- Decoder ----
        rev = tuple(reversed(ch))
        self.dec_in = _ConvW(ws, "decoder.conv_in", rev[0], latent_channels, 3)
        self.dec_mid = _MidW(ws, "decoder.mid_block", rev[0], mid_block_add_attention)
        self.dec_up = []
        c_prev = rev[0]
        for i, c in enumerate(rev):
            res = [_ResnetW(ws, f"decoder.up_blocks.{i}.resnets.{j}", c_prev if j == 0 else c, c) for j in range(layers_per_block + 1)]
            up = _ConvW(ws, f"decoder.up_blocks.{i}.upsamplers.0.conv", c, c, 3) if i < n - 1 else None
            self.dec_up.append((res, up))
            c_prev = c
        self.dec_norm_out = _NormW(ws, "decoder.conv_norm_out", rev[-1])
        self.dec_out = _ConvW(ws, "decoder.conv_out", out_channels, rev[-1], 3)
        self.conv_variant = 0
        self.max_score_elems = 1 << 30  # bf16 score elements per chunk of the mid-block attention (2 GB)

    _CONFIG_FIELDS = ("in_channels", "out_channels", "latent_channels", "block_out_channels", "layers_per_block", "norm_num_groups",
                      "scaling_factor", "shift_factor", "use_quant_conv", "use_post_quant_conv", "mid_block_add_attention")

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path: str, subfolder: Optional[str] = None, torch_dtype=None,
                        device: Any = "cuda", variant: Optional[str] = None, trust_pickle: bool = False, **kwargs):
        """`AutoencoderKL.from_pretrained(base, subfolder="vae")` — what the pipeline's own `from_pretrained(base, transformer=None)` does for its `vae` component (infer.py:146-149):
        reads `config.json` (diffusers AutoencoderKL fields) and `diffusion_pytorch_model*.safetensors` of a LOCAL folder. No hub
        download (this process has no network by design)."""
        import json
        import os
        from . import checkpoint
        path = os.path.join(pretrained_model_name_or_path, subfolder) if subfolder else pretrained_model_name_or_path
        cfg_file = os.path.join(path, "config.json")
        if not os.path.isfile(cfg_file):
            raise OSError(f"{path} does not contain config.json: from_pretrained needs a local diffusers VAE folder")
        with open(cfg_file) as f:
            cfg = json.load(f)
        if torch_dtype not in (None, BF16):
            import warnings
            warnings.warn(f"torch_dtype={torch_dtype}: the B200-native path stores weights and activations in bf16 (fp32 accumulate)")
        fields = {k: cfg[k] for k in cls._CONFIG_FIELDS if k in cfg}
        fields.update({k: v for k, v in kwargs.items() if k in cls._CONFIG_FIELDS})
        for k, want in (("act_fn", "silu"), ("down_block_types", "DownEncoderBlock2D"), ("up_block_types", "UpDecoderBlock2D")):
            got = cfg.get(k, want)
            if any(g != want for g in (got if isinstance(got, (list, tuple)) else [got])):
                raise ops.UgError(f"AutoencoderKL (B200-native): {k}={got!r} is not covered (only {want!r})")
        model = cls(device=device, **fields)
        sd = checkpoint.read_state_dict(path, trust_pickle=trust_pickle, variant=variant)
        own = model._ws.views
        # checkpoints converted from the original LDM layout keep the attention projections as 1x1 convolutions ([c, c, 1, 1])
        sd = {k: (v.reshape(v.shape[0], v.shape[1]) if k in own and v.dim() == 4 and own[k].dim() == 2 else v) for k, v in sd.items()}
        missing = [k for k in own if k not in sd]
        if missing:
            raise RuntimeError(f"{path} lacks {len(missing)} VAE keys, e.g. {missing[:4]}")
        model.load_state_dict({k: v for k, v in sd.items() if k in own}, strict=False)
        model.name_or_path = pretrained_model_name_or_path
        return model

    def to(self, *args, **kwargs):
        """`vae.to(device, dtype=...)`: the weights already live on the CUDA device the model was built on, in bf16; another
        device is an error (no fallback), a dtype is a no-op."""
        for a in list(args) + [kwargs.get("device")]:
            if isinstance(a, (str, torch.device)) and a is not None:
                d = torch.device(a)
                if d.type != "cuda" or (d.index is not None and d.index != (self.device_.index or 0)):
                    raise ops.UgError(f"the B200-native VAE lives on {self.device_}; moving it to {d} is not supported")
        return self

    def requires_grad_(self, requires_grad: bool = True):  # the pipelines keep the VAE frozen (train.py:312 `vae.requires_grad_(False)`)
        if requires_grad:
            raise ops.UgError("the B200-native path is forward-only")
        return self

    @torch.no_grad()
    def init_random_(self, seed: int = 0, zero_linear_std=None):
        """nn.Conv2d / nn.Linear default init on the device, GroupNorm affine = (1, 0) (bench / smoke: no checkpoint offline)."""
        super().init_random_(seed)
        for k, v in self._ws.views.items():
            if ".norm" in k or "group_norm" in k or "conv_norm_out" in k:
                v.fill_(1.0 if k.endswith(".weight") else 0.0)
        return self

    # ---------------------------------------------------------------------------------------------------------
    def _rec_nhwc(self, name: str, t: torch.Tensor, channels: Optional[int] = None):
        if self.trace is not None:
            self.trace[name] = ops.nhwc_to_nchw(t, channels or t.shape[3], dtype=torch.float32)

    def _conv(self, x: torch.Tensor, cw: _ConvW, residual: Optional[torch.Tensor] = None, stride: int = 1) -> torch.Tensor:
        """3x3 convolution over NHWC x (padding 1 for stride 1; `Downsample2D`'s (0, 1, 0, 1) padding for stride 2) or 1x1."""
        B, H, W, Ci = x.shape
        if cw.k == 1:
            out = torch.empty(B, H, W, cw.c_out_pad, device=x.device, dtype=BF16)
            ops.gemm(x.view(B, H * W, Ci), cw.w, out=out.view(B, H * W, cw.c_out_pad), bias=cw.b, variant=self.gemm_variant)
            return out
        if stride == 1 and ops.conv3x3_implicit_ok(Ci, W) and cw.k_pad == cw.k_cols:
            out = torch.empty(B, H, W, cw.c_out_pad, device=x.device, dtype=BF16)
            return ops.conv3x3(x, cw.w, bias=cw.b, residual=residual, out=out, variant=self.conv_variant)
        ho, wo = (H, W) if stride == 1 else (H // 2, W // 2)
        pad = 1 if stride == 1 else 0
        cols = ops.im2col(x, "nhwc", 3, 3, stride, pad, pad, ho, wo, cw.k_pad)
        return self._cols_gemm(cols, cw, B, ho, wo, residual)

    def _cols_gemm(self, cols: torch.Tensor, cw: _ConvW, B: int, ho: int, wo: int, residual: Optional[torch.Tensor] = None):
        out = torch.empty(B, ho, wo, cw.c_out_pad, device=cols.device, dtype=BF16)
        r = residual.view(B, ho * wo, residual.shape[3]) if residual is not None else None
        ops.gemm(cols.view(B, ho * wo, cw.k_pad), cw.w, out=out.view(B, ho * wo, cw.c_out_pad), bias=cw.b, residual=r,
                 variant=self.gemm_variant)
        return out

    def _resnet(self, x: torch.Tensor, w: _ResnetW) -> torch.Tensor:
        """ResnetBlock2D (temb None, output_scale_factor 1): the skip add rides in conv2's GEMM epilogue."""
        g = self.config.norm_num_groups
        h = self._conv(ops.groupnorm(x, w.norm1.w, w.norm1.b, g, silu_act=True), w.conv1)
        ops.groupnorm(h, w.norm2.w, w.norm2.b, g, silu_act=True, out=h)
        skip = self._conv(x, w.shortcut) if w.shortcut is not None else x
        return self._conv(h, w.conv2, residual=skip)

    def _attention(self, x: torch.Tensor, w: _AttnW) -> torch.Tensor:
        """Mid-block attention (one head of width C over the h * w pixels) + residual."""
        B, H, W, Cc = x.shape
        T, gv = H * W, self.gemm_variant
        t = ops.groupnorm(x, w.norm.w, w.norm.b, self.config.norm_num_groups, silu_act=False).view(B, T, Cc)
        q = ops.gemm(t, w.q[0], bias=w.q[1], variant=gv)
        k = ops.gemm(t, w.k[0], bias=w.k[1], variant=gv)
        if T % 64:
            raise ops.UgError(f"VAE attention needs h * w ({T}) to be a multiple of 64")
        o = torch.empty(B, T, Cc, device=x.device, dtype=BF16)
        vt = torch.empty(Cc, T, device=x.device, dtype=BF16)
        # query rows in chunks so that the bf16 score matrix stays <= ~2 GB (T = 16 384 at 1024^2: one 512 MB chunk)
        rows = min(T, max(64, (self.max_score_elems // T) // 64 * 64))
        s = torch.empty(rows, T, device=x.device, dtype=BF16)
        for b in range(B):
            # v^T [C, T] straight out of a GEMM (A = W_v, "weights" = the tokens); its bias joins after the softmax (rows of p sum to 1)
            ops.gemm(w.v[0].unsqueeze(0), t[b], out=vt.unsqueeze(0), variant=gv)
            for r0 in range(0, T, rows):
                n = min(rows, T - r0)
                ops.gemm(q[b:b + 1, r0:r0 + n], k[b], out=s[:n].unsqueeze(0), alpha=1.0 / math.sqrt(Cc), variant=gv)
                ops.softmax_rows_(s[:n])
                ops.gemm(s[:n].unsqueeze(0), vt, out=o[b:b + 1, r0:r0 + n], bias=w.v[1], variant=gv)
        out = torch.empty_like(x)
        ops.gemm(o, w.out[0], out=out.view(B, T, Cc), bias=w.out[1], residual=x.view(B, T, Cc), variant=gv)
        return out

    def _mid(self, x: torch.Tensor, w: _MidW) -> torch.Tensor:
        x = self._resnet(x, w.res0)
        if w.attn is not None:
            x = self._attention(x, w.attn)
        return self._resnet(x, w.res1)

    # ---------------------------------------------------------------------------------------------------------
    # reference API
    # ---------------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def _encode_moments(self, x: torch.Tensor) -> torch.Tensor:
        cfg = self.config
        if x.dim() != 4 or x.shape[1] != cfg.in_channels:
            raise ops.UgError(f"encode expects an image batch [B, {cfg.in_channels}, H, W], got {tuple(x.shape)}")
        x = x.to(self.device_)
        if x.dtype not in (BF16, torch.float32):
            x = x.float()
        B, _, H, W = x.shape
        f = 2 ** (len(cfg.block_out_channels) - 1)
        if H % f or W % f:
            raise ops.UgError(f"image size {H}x{W} must be a multiple of {f}")
        n0 = ops.launch_count()
        h = self._cols_gemm(ops.im2col(x, "nchw", 3, 3, 1, 1, 1, H, W, self.enc_in.k_pad), self.enc_in, B, H, W)
        self._rec_nhwc("encoder.conv_in", h)
        for i, (res, down) in enumerate(self.enc_down):
            for w in res:
                h = self._resnet(h, w)
            if down is not None:
                h = self._conv(h, down, stride=2)
            self._rec_nhwc(f"encoder.down_blocks.{i}", h)
        h = self._mid(h, self.enc_mid)
        self._rec_nhwc("encoder.mid_block", h)
        ops.groupnorm(h, self.enc_norm_out.w, self.enc_norm_out.b, cfg.norm_num_groups, silu_act=True, out=h)
        m = self._conv(h, self.enc_out)
        self._rec_nhwc("encoder.moments", m, 2 * cfg.latent_channels)
        self.last_launches = ops.launch_count() - n0
        return m

    def encode(self, x: torch.Tensor, return_dict: bool = True):
        """`vae.encode(x)` -> object with `.latent_dist` (`.sample(generator)`, `.mode()`): NCHW bf16 latents, unscaled."""
        dist = DiagonalGaussianDistribution(self._encode_moments(x), self.config.latent_channels)
        return types.SimpleNamespace(latent_dist=dist) if return_dict else (dist,)

    def encode_condition(self, x: torch.Tensor, generator: Optional[torch.Generator] = None, sample: bool = True,
                         use_shift_factor: bool = True) -> torch.Tensor:
        """The pipelines' three lines in one call (src/UniGenPipeline.py:306-308): `(vae.encode(x).latent_dist.sample() - shift) *
        scaling` — the shift / scale ride in the sampling kernel. `use_shift_factor` = the SD3 pipeline's `control_use_vae_shift_factor`."""
        dist = self.encode(x).latent_dist
        shift = self.config.shift_factor if use_shift_factor else 0.0
        return dist.sample(generator, shift, self.config.scaling_factor) if sample else dist.mode(shift, self.config.scaling_factor)

    @torch.no_grad()
    def decode(self, z: torch.Tensor, return_dict: bool = True, generator=None, _alpha: float = 1.0, _beta: float = 0.0,
               out_dtype=BF16):
        """`vae.decode(z)`: latents NCHW [B, latent_channels, h, w] -> image NCHW [B, out_channels, 8h, 8w] (bf16)."""
        cfg = self.config
        if z.dim() != 4 or z.shape[1] != cfg.latent_channels:
            raise ops.UgError(f"decode expects latents [B, {cfg.latent_channels}, h, w], got {tuple(z.shape)}")
        z = z.to(self.device_)
        if z.dtype not in (BF16, torch.float32):
            z = z.float()
        B, _, H, W = z.shape
        n0 = ops.launch_count()
        cols = ops.im2col(z, "nchw", 3, 3, 1, 1, 1, H, W, self.dec_in.k_pad, alpha=_alpha, beta=_beta)
        h = self._cols_gemm(cols, self.dec_in, B, H, W)
        self._rec_nhwc("decoder.conv_in", h)
        h = self._mid(h, self.dec_mid)
        self._rec_nhwc("decoder.mid_block", h)
        for i, (res, up) in enumerate(self.dec_up):
            for w in res:
                h = self._resnet(h, w)
            if up is not None:
                h = self._conv(ops.upsample2x(h), up)
            self._rec_nhwc(f"decoder.up_blocks.{i}", h)
        ops.groupnorm(h, self.dec_norm_out.w, self.dec_norm_out.b, cfg.norm_num_groups, silu_act=True, out=h)
        img = self._conv(h, self.dec_out)
        sample = ops.nhwc_to_nchw(img, cfg.out_channels, dtype=out_dtype)
        if self.trace is not None:
            self.trace["decoder.sample"] = sample.float()
        self.last_launches = ops.launch_count() - n0
        return types.SimpleNamespace(sample=sample) if return_dict else (sample,)

    def decode_latents(self, latents: torch.Tensor, out_dtype=BF16) -> torch.Tensor:
        """The pipelines' two lines in one call (src/UniGenPipeline.py:439-441): `vae.decode(latents / scaling_factor + shift_factor)`
        — the affine rides in the first convolution's patch gather (padding stays zero, as in the reference)."""
        return self.decode(latents, return_dict=False, _alpha=1.0 / self.config.scaling_factor, _beta=self.config.shift_factor,
                           out_dtype=out_dtype)[0]

    def forward(self, sample: torch.Tensor, sample_posterior: bool = False, generator=None):
        dist = self.encode(sample).latent_dist
        z = dist.sample(generator) if sample_posterior else dist.mode()
        return self.decode(z)
