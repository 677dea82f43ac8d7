"""Denoise-loop glue on the device vs the oracle loop."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_pack_unpack_cfg_euler_kernels(ug):
    from oracle import unigen_oracle as O
    torch.manual_seed(0)
    x = torch.randn(2, 16, 12, 20).to(torch.bfloat16)
    p = ug.pack_latents(x.cuda())
    assert torch.equal(p.cpu(), O.pack_latents(x))                      # index permutation: bit-exact
    assert torch.equal(ug.unpack_latents(p, 12, 20).cpu(), x)
    assert torch.equal(O.unpack_latents(O.pack_latents(x), 12, 20), x)
    u, t = torch.randn(1000).to(torch.bfloat16), torch.randn(1000).to(torch.bfloat16)
    got = ug.cfg_combine(u.cuda(), t.cuda(), 3.5).cpu()
    assert torch.equal(got, (u.float() + 3.5 * (t.float() - u.float())).to(torch.bfloat16))
    lat, v = torch.randn(4096).to(torch.bfloat16), torch.randn(4096).to(torch.bfloat16)
    want = (lat.float() + (0.25 - 0.75) * v.float()).to(torch.bfloat16)
    assert torch.equal(ug.euler_step(lat.cuda().clone(), v.cuda(), 0.75, 0.25).cpu(), want)


def test_sigma_schedule_matches_oracle():
    from oracle import unigen_oracle as O
    from unigen_b200.pipeline import flow_match_sigmas
    for n, seq in [(4, 256), (4, 4096), (28, 1024), (1, 256)]:
        assert torch.allclose(torch.tensor(flow_match_sigmas(n, seq)), O.flow_match_sigmas(n, seq), atol=1e-6)


def test_four_step_denoise_matches_oracle_loop():
    from oracle import unigen_oracle as O
    from unigen_b200.model import FluxArch, UniGenFlux, canonical_control_params
    from unigen_b200.pipeline import denoise
    cfg = O.FluxConfig.tiny()
    sd = O.init_state_dict(cfg, seed=0)
    sd = {k: (v if k.endswith("gate.wg.weight") else v.to(torch.bfloat16).float()) for k, v in sd.items()}
    inp = O.make_inputs(cfg, 256, 256, text_len=512)
    for k in ("hidden_states", "condition_hidden_states", "encoder_hidden_states"):
        inp[k] = inp[k].to(torch.bfloat16).float()
    g = torch.Generator().manual_seed(11)
    rts = [torch.rand(256, cfg.expert_nums, generator=g) for _ in range(4)]
    base = {k: v for k, v in inp.items() if k not in ("rts_uniform", "timestep")}
    want = O.denoise_loop(O.UniGenFluxOracle(cfg, sd), dict(base), 4, rts)
    model = UniGenFlux(FluxArch.tiny(), device="cuda")
    model.init_condition_block(condition_nums=1, control_params=canonical_control_params())
    model.load_state_dict(sd)
    cu = lambda t: t.cuda()  # noqa: E731
    got = denoise(model, cu(inp["hidden_states"]), cu(inp["condition_hidden_states"]), cu(inp["encoder_hidden_states"]),
                  cu(inp["pooled_projections"]), cu(inp["condition_pooled_projections"]), cu(inp["img_ids"]), cu(inp["txt_ids"]),
                  cu(inp["condition_ids"]), num_inference_steps=4, rts_uniform=[cu(r) for r in rts])
    got = got.float().cpu()
    rel = ((got - want).norm() / want.norm()).item()
    cos = torch.nn.functional.cosine_similarity(got.flatten(), want.flatten(), dim=0).item()
    assert rel < 1e-2 and cos >= 0.999, (rel, cos)
