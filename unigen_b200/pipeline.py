"""Denoise loop around the B200-native transformer: the part of `UniGenFLUXPipeline.__call__`
between latent preparation and VAE decode (reference src/UniGenPipeline.py:989-1006 sigma schedule, :1050-1116 loop body),
with the scheduler update kept on the device. Text encoders, VAE and image preprocessing stay out of scope."""
from __future__ import annotations

import math
from typing import List, Optional, Sequence

import torch

from . import ops


def calculate_shift(image_seq_len: int, base_seq_len: int = 256, max_seq_len: int = 4096, base_shift: float = 0.5,
                    max_shift: float = 1.16) -> float:
    """diffusers pipeline_flux.calculate_shift (call site src/UniGenPipeline.py:991-997)."""
    m = (max_shift - base_shift) / (max_seq_len - base_seq_len)
    b = base_shift - m * base_seq_len
    return image_seq_len * m + b


def flow_match_sigmas(num_inference_steps: int, image_seq_len: int, use_dynamic_shifting: bool = True) -> List[float]:
    """sigmas = linspace(1, 1/n, n) (src/UniGenPipeline.py:989) -> FlowMatchEulerDiscreteScheduler.set_timesteps with
    dynamic (exponential) time shifting mu = calculate_shift(seq_len); a terminal 0 is appended. timestep_i = 1000 sigma_i."""
    n = num_inference_steps
    sig = [1.0 + (1.0 / n - 1.0) * i / max(n - 1, 1) for i in range(n)]
    if use_dynamic_shifting:
        mu = calculate_shift(image_seq_len)
        sig = [math.exp(mu) / (math.exp(mu) + (1.0 / s - 1.0)) for s in sig]
    return sig + [0.0]


@torch.no_grad()
def denoise(transformer, latents: torch.Tensor, condition_latents, encoder_hidden_states, pooled_projections,
            condition_pooled_projections, img_ids, txt_ids, condition_ids, num_inference_steps: int = 4,
            guidance: Optional[torch.Tensor] = None, conditioning_scale: float = 1.0, use_dynamic_shifting: bool = True,
            rts_uniform: Optional[Sequence] = None) -> torch.Tensor:
    """Runs the sampling loop on packed latents (B, N, 64) and returns the final packed latents (bf16, on the device).
    Per step: velocity = transformer(latents, ..., timestep = sigma_i)[0]; latents += (sigma_{i+1} - sigma_i) * velocity."""
    dev = transformer.device
    x = ops.to_bf16(latents.to(dev).contiguous()).clone()
    sig = flow_match_sigmas(num_inference_steps, x.shape[1], use_dynamic_shifting)
    B = x.shape[0]
    for i in range(num_inference_steps):
        t = torch.full((B,), sig[i], device=dev, dtype=torch.float32)  # pipeline passes timestep / 1000 == sigma
        v = transformer(hidden_states=x, condition_hidden_states=condition_latents, conditioning_scale=conditioning_scale,
                        encoder_hidden_states=encoder_hidden_states, pooled_projections=pooled_projections,
                        condition_pooled_projections=condition_pooled_projections, timestep=t, img_ids=img_ids, txt_ids=txt_ids,
                        guidance=guidance, condition_ids=condition_ids,
                        rts_uniform=None if rts_uniform is None else rts_uniform[i])[0]
        ops.euler_step(x, v.contiguous(), sig[i], sig[i + 1])
    return x
