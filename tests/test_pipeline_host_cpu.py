"""CPU: host-side schedule logic of unigen_b200/pipeline.py (no kernels involved) against the oracle's restatement of the reference
pipelines — sigma schedules (src/UniGenPipeline.py:989-1006), the control-guidance window (:367-373) — and the VAE weight layout
(views with nn.Conv2d's shapes over the GEMM matrices)."""
import math

import pytest
import torch

from oracle import unigen_oracle as O
from unigen_b200 import pipeline as PL


@pytest.mark.parametrize("steps,seq", [(4, 1024), (28, 4096), (50, 256), (1, 4096)])
def test_flux_sigma_schedule_matches_the_oracle(steps, seq):
    want = O.flow_match_sigmas(steps, seq)
    got = PL.flow_match_sigmas(steps, seq)
    assert len(got) == steps + 1 and got[-1] == 0.0
    torch.testing.assert_close(torch.tensor(got, dtype=torch.float64), want.double(), rtol=1e-6, atol=1e-7)
    assert math.isclose(PL.calculate_shift(seq), O.calculate_shift(seq), rel_tol=1e-12)


def test_sd3_static_shift_schedule():
    """FlowMatchEulerDiscreteScheduler(shift=3.0, use_dynamic_shifting=False) as the SD3 pipeline drives it (no `sigmas` argument,
    src/UniGenPipeline.py:345-351): __init__ shifts the training schedule linspace(1, 1000, 1000) / 1000 once (sigma_max = 1,
    sigma_min = shift(1 / 1000)); set_timesteps takes linspace(sigma_max, sigma_min, n) and shifts it again; terminal 0."""
    n, shift = 28, 3.0
    f = lambda s: shift * s / (1 + (shift - 1) * s)  # noqa: E731
    train = f(torch.linspace(1.0, 1000.0, 1000, dtype=torch.float64).flip(0) / 1000.0)
    s = torch.linspace(train[0].item(), train[-1].item(), n, dtype=torch.float64)
    want = torch.cat([f(s), torch.zeros(1, dtype=torch.float64)])
    got = PL.flow_match_sigmas(n, 4096, use_dynamic_shifting=False, shift=shift)
    torch.testing.assert_close(torch.tensor(got, dtype=torch.float64), want, rtol=1e-9, atol=1e-12)
    assert got[0] == 1.0 and abs(got[n - 1] - f(0.003 / 1.002)) < 1e-12
    # explicit sigmas (the `sigmas=` argument of the pipeline) are shifted once
    torch.testing.assert_close(torch.tensor(PL.flow_match_sigmas(2, 1, False, sigmas=[1.0, 0.5], shift=shift)),
                               torch.tensor([1.0, 0.75, 0.0]))


@pytest.mark.parametrize("n,start,end", [(4, 0.0, 1.0), (4, 0.0, 0.5), (4, 0.25, 1.0), (28, 0.1, 0.8), (7, 0.0, 0.0), (3, 0.5, 0.5)])
def test_controlnet_keep_is_the_reference_expression(n, start, end):
    """src/UniGenPipeline.py:367-373."""
    want = [1.0 - float(i / n < start or (i + 1) / n > end) for i in range(n)]
    assert [PL.sd3_controlnet_keep(i, n, start, end) for i in range(n)] == want


def test_vae_conv_weight_views_have_conv2d_shapes_over_the_gemm_layout():
    from unigen_b200.model import _Weights
    from unigen_b200.vae import _ConvW
    ws = _Weights("cpu")
    cw = _ConvW(ws, "c", 3, 16, 3)  # 16 -> 3 channels: both GEMM dimensions padded to one 64-element TMA box row
    assert cw.w.shape == (64, 192) and cw.k_cols == 144 and cw.c_out_pad == 64
    w = torch.randn(3, 16, 3, 3)
    ws.views["c.weight"].copy_(w)
    ws.views["c.bias"].copy_(torch.arange(3.0))
    assert ws.views["c.weight"].shape == (3, 16, 3, 3)
    # column (ky * 3 + kx) * c_in + ci of the GEMM matrix holds weight[co, ci, ky, kx]; padding stays zero
    assert torch.equal(cw.w[:3, :144].float(), w.permute(0, 2, 3, 1).reshape(3, 144).to(torch.bfloat16).float())
    assert (cw.w[3:] == 0).all() and (cw.w[:, 144:] == 0).all() and (cw.b[3:] == 0).all()


def test_condition_encode_image_builds_tokens_and_ids_like_the_reference():
    """`Condition._encode_image` host logic (src/condition.py:90-111) with a stand-in pipe: VAE latents (already shifted / scaled) ->
    packed tokens; ids of the latent grid // 2; `subject` offsets column 2 by the latent height // 2; type ids from condition_dict."""
    import types
    from unigen_b200.condition import Condition
    lat = torch.arange(2 * 16 * 8 * 12, dtype=torch.float32).reshape(2, 16, 8, 12)
    calls = {}

    def encode_condition(img, generator=None):
        calls["img"], calls["gen"] = img, generator
        return lat

    def pack(x):  # FluxPipeline._pack_latents
        B, C, H, W = x.shape
        return x.view(B, C, H // 2, 2, W // 2, 2).permute(0, 2, 4, 1, 3, 5).reshape(B, (H // 2) * (W // 2), C * 4)

    pipe = types.SimpleNamespace(vae=types.SimpleNamespace(encode_condition=encode_condition), _pack_latents=pack)
    img = torch.zeros(2, 3, 64, 96)
    gen = torch.Generator().manual_seed(1)
    tokens, ids, type_id = Condition("depth", img).encode(pipe, gen)
    assert calls["img"] is img and calls["gen"] is gen
    assert tokens.shape == (2, 24, 64) and torch.equal(tokens, O.pack_latents(lat))
    assert torch.equal(ids, O.prepare_latent_image_ids(4, 6)) and torch.equal(type_id, torch.zeros(24, 1))
    _, sub_ids, sub_type = Condition("subject", img).encode(pipe)
    assert torch.equal(sub_ids[:, 2], ids[:, 2] + 4) and torch.equal(sub_ids[:, :2], ids[:, :2]) and int(sub_type[0, 0]) == 4
    with pytest.raises(ValueError):
        Condition("canny", img).encode(types.SimpleNamespace(vae=None))


def test_sd3_modulated_experts_equal_the_flux_oracle_expert_path():
    """Two independent restatements of the same reference arithmetic (`UniGenBase.expert_forward` :252-255 and `UniGenFlux.expert_forward`
    :953-959 are the same lines): the SD3 oracle's `use_modulate` branch and the Flux oracle's expert path agree on shared weights."""
    import dataclasses
    from oracle import unigen_sd3_oracle as S
    g = torch.Generator().manual_seed(5)
    cfg3 = dataclasses.replace(S.SD3Config.tiny(), use_modulate=True, num_attention_heads=2, attention_head_dim=16, pooled_projection_dim=24,
                               condition_nums=0, expert_num_each_condition=3)
    D, E, C, P = cfg3.inner_dim, cfg3.expert_nums, 5, 24
    sd = {}
    for e in range(E):
        for br in (0, 1):
            p = f"moe.moe_layer.experts.deepspeed_experts.{e}.{br}"
            sd[p + ".0.weight"], sd[p + ".0.bias"] = torch.randn(D, D, generator=g) / D ** 0.5, torch.randn(D, generator=g)
            sd[p + ".1.weight"], sd[p + ".1.bias"] = torch.randn(D, P, generator=g) / P ** 0.5, torch.randn(D, generator=g)
    hidden, cond = torch.randn(1, E, C, D, generator=g), torch.randn(1, E, C, D, generator=g)
    pooled, cpooled = torch.randn(1, E, C, P, generator=g), torch.randn(1, E, C, P, generator=g)
    eh3, ec3 = S.UniGenSD3Oracle(cfg3, sd).expert_forward(hidden, cond, None, None, pooled, cpooled)
    fcfg = dataclasses.replace(O.FluxConfig.tiny(), condition_nums=0, expert_num_each_condition=3)
    assert fcfg.expert_nums == E
    ehf, ecf = O.UniGenFluxOracle(fcfg, sd).expert_forward(hidden, cond, pooled, cpooled)
    torch.testing.assert_close(eh3, ehf)
    torch.testing.assert_close(ec3, ecf)
