"""The whole-step C handle (`ug_flux_create / ug_flux_bind_weight / ug_flux_workspace_bytes / ug_flux_forward`):
a step sequenced entirely inside libunigen_b200.so is bit-identical to the step the Python mirror sequences (same kernels, same
order), reports the same routing statistics, replays under CUDA-graph capture, and fails loudly on unbound / mis-laid-out weights."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _setup(n_cond=1, **cfg_kw):
    from oracle import unigen_oracle as O
    from unigen_b200.model import FluxArch, UniGenFlux, canonical_control_params
    cfg = O.FluxConfig.tiny()
    cfg.condition_nums = n_cond
    for k, v in cfg_kw.items():
        setattr(cfg, k, v)
    sd = O.init_state_dict(cfg, seed=4)
    sd = {k: (v if k.endswith("gate.wg.weight") else v.to(torch.bfloat16).float()) for k, v in sd.items()}
    inp = O.make_multi_inputs(cfg, 256, 256) if n_cond > 1 else O.make_inputs(cfg, 320, 192, text_len=77, batch=2)
    model = UniGenFlux(FluxArch(num_layers=2, num_single_layers=4, attention_head_dim=64, num_attention_heads=6, axes_dims_rope=(8, 28, 28),
                                guidance_embeds=cfg.guidance_embeds), device="cuda")
    params = dict(canonical_control_params(), single_block_control_method=cfg.single_block_control_method)
    model.init_condition_block(condition_nums=n_cond, control_params=params)
    model.load_state_dict(sd, strict=True)
    to_dev = lambda v: [t.cuda() for t in v] if isinstance(v, list) else (v.cuda() if torch.is_tensor(v) else v)  # noqa: E731
    return cfg, {k: to_dev(v) for k, v in inp.items()}, model


@pytest.mark.parametrize("variant", ["single", "multi", "guidance_single_add"])
def test_c_handle_step_is_bit_identical_to_the_python_sequenced_step(variant):
    from unigen_b200.chandle import FluxStepHandle
    if variant == "multi":
        cfg, inp, model = _setup(3)
    elif variant == "guidance_single_add":
        cfg, inp, model = _setup(1, guidance_embeds=True, single_block_control_method="single_add")
        inp["guidance"] = torch.tensor([3.5, 1.0]).cuda()
        inp["conditioning_scale"] = 0.6
    else:
        cfg, inp, model = _setup(1)
    want, want_l, want_o = model(**inp)
    h = FluxStepHandle(model)
    got, moe_loss, counts = h.forward(**inp)
    torch.cuda.synchronize()
    assert torch.equal(got, want)
    assert torch.equal(counts, want_o["expert_counts"]) and moe_loss.item() == want_l["moe_loss"].item()
    # a second call reuses the cached job table; capture + replay of the call is allowed from then on
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        h.forward(**inp)
    torch.cuda.current_stream().wait_stream(s)
    with torch.cuda.graph(g):
        again = h.forward(**inp)[0]
    g.replay()
    torch.cuda.synchronize()
    assert torch.equal(again, want)
    h.close()


def test_c_handle_reports_unbound_and_non_contiguous_weights():
    from unigen_b200 import _lib
    from unigen_b200.chandle import FluxStepHandle
    from unigen_b200.ops import UgError
    cfg, inp, model = _setup(1)
    h = FluxStepHandle(model)
    # re-bind to_k somewhere else: q | k | v no longer form one contiguous block
    name = "transformer_blocks.0.attn.to_k.weight"
    stray = model.state_dict()[name].clone()
    h.bind_state_dict({name: stray})
    with pytest.raises(UgError, match="contiguous"):
        h.forward(**inp)
    h.bind_state_dict({name: model.state_dict()[name]})
    assert torch.equal(h.forward(**inp)[0], model(**inp)[0])
    h.close()
    # a wrong shape / dtype is named, not silently accepted
    h2 = FluxStepHandle(model)
    h2.bind_state_dict({"proj_out.weight": torch.zeros(64, 128, device="cuda", dtype=torch.bfloat16)})
    with pytest.raises(UgError, match="proj_out.weight"):
        h2.forward(**inp)
    h2.bind_state_dict({"proj_out.weight": model.state_dict()["proj_out.weight"],
                        "moe.moe_layer.gate.wg.weight": model.state_dict()["moe.moe_layer.gate.wg.weight"].to(torch.bfloat16)})
    with pytest.raises(UgError, match="gate.wg.weight"):
        h2.forward(**inp)
    h2.close()
