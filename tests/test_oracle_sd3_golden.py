"""CPU: the SD3.5 oracle restatement (oracle/unigen_sd3_oracle.py) replayed against golden vectors produced by the
REAL reference functions (tests/golden/make_golden_sd3.py)."""
from pathlib import Path

import pytest
import torch

from oracle import unigen_sd3_oracle as O
from oracle.unigen_oracle import linear

GOLD = Path(__file__).parent / "golden" / "reference_golden_sd3.pt"
TOL = dict(rtol=1e-5, atol=1e-5)


@pytest.fixture(scope="module")
def gold():
    return torch.load(GOLD, weights_only=False)


def _pref(w, p):
    return {f"{p}.{k}": v for k, v in w.items()}


def test_adaln_forwards(gold):
    for c in gold["adaln"]:
        sd = _pref(c["w"], "m")
        if c["kind"] == "zero":
            out = O.ada_norm_zero(sd, "m", c["x"], c["emb"])
        elif c["kind"] == "zero_x":
            out = O.ada_norm_zero_x(sd, "m", c["x"], c["emb"])
        else:
            out = (O.ada_norm_continuous(sd, "m", c["x"], c["emb"]),)
        assert len(out) == len(c["out"])
        for a, b in zip(out, c["out"]):
            torch.testing.assert_close(a, b, **TOL)


def test_block_forwards(gold):
    H = gold["heads"]
    seen = set()
    for c in gold["blocks"]:
        sd = _pref(c["w"], "b")
        if c["kind"] == "joint":
            enc, h = O.joint_block(sd, "b", H, c["h"], c["c"], c["temb"], c["dual"], c["cpo"])
            torch.testing.assert_close(h, c["h_out"], **TOL)
            assert (enc is None) == (c["enc_out"] is None)
            if enc is not None:
                torch.testing.assert_close(enc, c["enc_out"], **TOL)
            seen.add(("joint", c["dual"], c["cpo"], c["temb"].dim()))
        else:
            torch.testing.assert_close(O.sd3_single_block(sd, "b", H, c["x"], c["temb"]), c["y"], **TOL)
            seen.add(("single", c["temb"].dim()))
    assert {("joint", True, True, 2), ("joint", False, False, 3), ("single", 3)} <= seen


def test_moe_with_transformer_block_experts_and_shared_experts(gold):
    m = gold["moe"]
    E = m["E"]
    D = m["hidden"].shape[-1]
    cfg = O.SD3Config(num_attention_heads=gold["heads"], attention_head_dim=D // gold["heads"], condition_nums=0,
                      expert_num_each_condition=E)
    assert cfg.expert_nums == E
    sd = {"moe.moe_layer.gate.wg.weight": m["wg"]}
    for e, pair in enumerate(m["experts"]):
        for br, w in enumerate(pair):
            sd.update(_pref(w, f"moe.moe_layer.experts.deepspeed_experts.{e}.{br}"))
    for s, w in enumerate(m["shared"]):
        sd.update(_pref(w, f"shared_expert.{s}"))
    orc = O.UniGenSD3Oracle(cfg, sd)
    eh, ec, l_aux, counts = orc.moe_forward(m["hidden"], m["cond"], m["enc"], m["temb"], m["ctemb"], m["rts"])
    torch.testing.assert_close(eh, m["out_hidden"], **TOL)
    torch.testing.assert_close(ec, m["out_cond"], **TOL)
    torch.testing.assert_close(l_aux, m["l_aux"])
    assert torch.equal(counts, m["counts"])


def test_moe_with_modulated_linear_experts_with_and_without_shared_experts(gold):
    """use_modulate=True (src/UniGenTransformer.py:171-183, :252-255) x use_shared_expert True / False (:279)."""
    assert [c["use_shared_expert"] for c in gold["moe_modulate"]] == [True, False]
    for m in gold["moe_modulate"]:
        E, D = m["E"], m["hidden"].shape[-1]
        cfg = O.SD3Config(num_attention_heads=gold["heads"], attention_head_dim=D // gold["heads"], condition_nums=0,
                          expert_num_each_condition=E, use_modulate=True, use_shared_expert=m["use_shared_expert"],
                          pooled_projection_dim=m["P"])
        sd = {"moe.moe_layer.gate.wg.weight": m["wg"]}
        for e, pair in enumerate(m["experts"]):
            for br, w in enumerate(pair):
                sd.update(_pref(w, f"moe.moe_layer.experts.deepspeed_experts.{e}.{br}"))
        for s, w in enumerate(m["shared"]):
            sd.update(_pref(w, f"shared_expert.{s}"))
        orc = O.UniGenSD3Oracle(cfg, sd)
        eh, ec, l_aux, counts = orc.moe_forward(m["hidden"], m["cond"], m["enc"], m["temb"], m["ctemb"], m["rts"],
                                                m["pooled"], m["cpooled"])
        torch.testing.assert_close(eh, m["out_hidden"], **TOL)
        torch.testing.assert_close(ec, m["out_cond"], **TOL)
        torch.testing.assert_close(l_aux, m["l_aux"])
        assert torch.equal(counts, m["counts"])


def test_weave_matches_reference_call_order_and_values(gold):
    """UniGenSD3.base_forward / control_forward / preprocess_moe_forward (src/UniGenTransformer.py:498-623)."""
    for w in gold["weave"]:
        n_base, n_ctrl = w["n_base"], w["n_ctrl"]
        calls = []

        def blk(tag, i, last, h, enc, temb):
            calls.append((tag, i))
            a, b = 1.0 + 0.01 * (i + 1), 0.1 * (i + 1)
            enc_o = None if last else enc * a + temb[:, None] * 0.01
            return enc_o, h * a + b + enc.mean(1, keepdim=True) * 0.05 + temb[:, None] * 0.02

        Dm = w["h0"].shape[-1]
        h, enc = w["h0"], w["e0"]
        moe = None
        for i in range(n_base):
            enc_prev = enc
            enc, h = blk("base", i, i == n_base - 1, h, enc, w["temb"])
            j = int(i / (n_base / n_ctrl))
            if i == 0:
                cond = w["cond_lat"].flatten(2).transpose(1, 2)[..., :Dm] * 2.0 + 0.3
                control_temb = w["timestep"][:, None] * 0.001 + w["pooled"][:, :Dm]
                condition_temb = w["timestep"][:, None] * 0.002 - w["cpooled"][:, :Dm]
                enc_ctrl = enc * 1.5 - 0.2
                kw = w["moe_kwargs"]  # what the reference handed to the MoE
                torch.testing.assert_close(kw["hidden_states"], h)
                torch.testing.assert_close(kw["condition_hidden_states"], cond)
                torch.testing.assert_close(kw["encoder_hidden_states"], enc_ctrl)
                torch.testing.assert_close(kw["temb"], control_temb)
                torch.testing.assert_close(kw["condition_temb"], condition_temb)
                torch.testing.assert_close(kw["pooled_projections"], w["pooled"])
                torch.testing.assert_close(kw["condition_pooled_projections"], w["cpooled"])
                moe = dict(enc=enc_ctrl, temb=condition_temb)
                ctrl_in = (h * 0.5 + 0.1) + (cond * 0.25 - 0.1)
            else:
                ctrl_in = h
            _, ch = blk("ctrl", j + 50, False, ctrl_in, moe["enc"], moe["temb"])
            h = h + linear({"a.weight": w["adders"][j]["weight"], "a.bias": w["adders"][j]["bias"]}, "a", ch) * 0.7
            del enc_prev
        assert calls == [tuple(c) for c in w["calls"]]
        torch.testing.assert_close(h, w["out_hidden"], **TOL)


def test_unpatchify(gold):
    u = gold["unpatchify"]
    assert torch.equal(O.unpatchify(u["tokens"], u["h"], u["w"], u["p"], u["c"]), u["out"])
    assert float(u["moe_loss"]) == pytest.approx(0.2)  # moe_loss * 0.1 (:680)


def test_tiny_forward_runs_and_is_deterministic():
    cfg = O.SD3Config.tiny()
    sd = O.init_state_dict(cfg)
    inp = O.make_inputs(cfg, 256, 256, text_len=77)
    a = O.UniGenSD3Oracle(cfg, sd).forward(**inp)[0]
    b = O.UniGenSD3Oracle(cfg, sd).forward(**inp)[0]
    assert a.shape == (1, 16, 32, 32) and torch.equal(a, b) and torch.isfinite(a).all()
    # true zero-linears: the control branch contributes exactly 0 -> bare SD3 base model
    sd0 = O.init_state_dict(cfg, zero_linear_std=None)
    full = O.UniGenSD3Oracle(cfg, sd0).forward(**inp)[0]
    H = cfg.num_attention_heads
    h = O.patch_embed(sd0, "pos_embed", inp["hidden_states"], cfg)
    temb = O.combined_timestep_text_embed(sd0, "time_text_embed", inp["timestep"], inp["pooled_projections"])
    enc = linear(sd0, "context_embedder", inp["encoder_hidden_states"])
    for i in range(cfg.num_layers):
        enc, h = O.joint_block(sd0, f"transformer_blocks.{i}", H, h, enc, temb, i in cfg.dual_attention_layers,
                               i == cfg.num_layers - 1)
    base = O.unpatchify(linear(sd0, "proj_out", O.ada_norm_continuous(sd0, "norm_out", h, temb)), 16, 16, 2, 16)
    assert torch.equal(full, base)
