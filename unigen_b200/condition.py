"""Token-index construction for condition inputs — the integer part of the reference's `Condition`
(src/condition.py:12-19 `condition_dict`, :101-111 `_encode_image` ids, :134 `type_id`) and of
`FluxPipeline._prepare_latent_image_ids` (call sites src/condition.py:101-108, src/UniGenPipeline.py:640-647).

BIT-EXACT integer work (north_star "bit-exact ... token-index construction"): the ids are small integers stored in the
pipeline dtype; every value here is exactly representable in bf16 / fp16 / fp32 up to 256 rows / columns per axis offset
(a 4096^2 image). Image pre-processing (canny / depth / blur) and the VAE encode of the condition image stay with the caller:
`Condition` here carries already-encoded packed latents, the `condition=` / `condition_ids=` form of the reference class."""
from __future__ import annotations

from typing import Optional, Tuple

import torch

# src/condition.py:12-19
condition_dict = {"depth": 0, "canny": 1, "subject": 4, "coloring": 6, "deblurring": 7, "fill": 9}


def prepare_latent_image_ids(height: int, width: int, device=None, dtype=torch.float32) -> torch.Tensor:
    """FluxPipeline._prepare_latent_image_ids(batch, height, width, device, dtype) with height / width the PACKED grid
    (latent // 2): (height * width, 3) rows (0, row, col)."""
    ids = torch.zeros(height, width, 3)
    ids[..., 1] = ids[..., 1] + torch.arange(height)[:, None]
    ids[..., 2] = ids[..., 2] + torch.arange(width)[None, :]
    return ids.reshape(height * width, 3).to(device=device, dtype=dtype)


def condition_ids(condition_type: str, height: int, width: int, device=None, dtype=torch.float32) -> Tuple[torch.Tensor, torch.Tensor]:
    """(ids, type_id) of a condition image of `height` x `width` PIXELS (VAE factor 8, 2x2 packing -> //16 per axis):
    ids = _prepare_latent_image_ids(latent_h // 2, latent_w // 2); `subject` shifts the column ids by latent_h // 2
    (`cond_ids[:, 2] += cond_img.shape[2] // 2`, src/condition.py:109-110 — the HEIGHT, as written there);
    type_id = ones_like(ids[:, :1]) * condition_dict[type] (:134)."""
    if condition_type not in condition_dict:
        raise NotImplementedError(f"unknown condition type {condition_type!r} (src/condition.py:118-132)")
    h2, w2 = height // 16, width // 16
    ids = prepare_latent_image_ids(h2, w2)
    if condition_type == "subject":
        ids[:, 2] += h2
    type_id = torch.ones_like(ids[:, :1]) * condition_dict[condition_type]
    return ids.to(device=device, dtype=dtype), type_id.to(device=device, dtype=dtype)


class Condition:
    """The `condition=` / `condition_ids=` form of the reference `Condition` (src/condition.py:22-48): packed condition
    latents (B, Nc, 64) plus their ids; `encode()` returns `(tokens, ids, type_id)` like the reference (:113-135). When no ids
    are given they are built from the pixel size of the condition image."""

    def __init__(self, condition_type: str, condition: torch.Tensor, condition_ids: Optional[torch.Tensor] = None,
                 height: Optional[int] = None, width: Optional[int] = None, mask=None):
        assert mask is None, "Mask not supported yet"  # src/condition.py:47
        if condition_type not in condition_dict:
            raise NotImplementedError(f"unknown condition type {condition_type!r}")
        self.condition_type = condition_type
        self.condition = condition
        self.condition_ids = condition_ids
        self.height, self.width = height, width

    @property
    def type_id(self) -> int:
        return condition_dict[self.condition_type]

    def _encode_image(self, pipe, cond_img: torch.Tensor, generator=None):
        """src/condition.py:90-111: VAE-encode the (already preprocessed, [-1, 1]) condition image, shift / scale, pack into
        tokens, build the position ids of the latent grid (`subject`: column 2 offset by the latent height // 2)."""
        vae = getattr(pipe, "vae", None)
        if vae is None:
            raise ValueError("Condition holds a pixel image: the pipeline needs a `vae` (unigen_b200.vae.AutoencoderKL) to encode it")
        lat = vae.encode_condition(cond_img, generator=generator)
        tokens = pipe._pack_latents(lat)
        ids = prepare_latent_image_ids(lat.shape[2] // 2, lat.shape[3] // 2, device=lat.device, dtype=torch.float32)
        if self.condition_type == "subject":
            ids[:, 2] += lat.shape[2] // 2
        return tokens, ids

    def encode(self, pipe=None, generator=None):
        tokens, ids = self.condition, self.condition_ids
        if ids is None and tokens.dim() == 4:  # a pixel image (B, 3, H, W), not packed tokens (B, Nc, 64): way (2) of the reference
            tokens, ids = self._encode_image(pipe, tokens, generator)
        if ids is None:
            if self.height is None or self.width is None:
                raise ValueError("Condition needs either condition_ids or the pixel height / width of the condition image")
            ids, _ = condition_ids(self.condition_type, self.height, self.width, device=tokens.device, dtype=tokens.dtype)
            if ids.shape[0] != tokens.shape[-2]:
                raise ValueError(f"{tokens.shape[-2]} condition tokens do not match a {self.height}x{self.width} condition image")
        type_id = torch.ones_like(ids[:, :1]) * self.type_id
        return tokens, ids, type_id
