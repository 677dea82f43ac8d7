"""Pins the oracle against golden vectors produced by the REAL reference functions (tests/golden/make_golden.py):
modulated_flatten, MOELayer dispatch/combine + expert_forward, the weave of base_forward/control_forward,
enable_lora, Condition ids. CPU only."""
import types

import torch

from oracle import unigen_oracle as O


def test_modulated_flatten_both_branches(golden):
    g = golden["modflat"]
    torch.testing.assert_close(O.modulated_flatten(g["x"], g["w"], g["s2"]), g["y2"], rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(O.modulated_flatten(g["x"], g["w"], g["s3"]), g["y3"], rtol=1e-5, atol=1e-5)


def test_moe_dispatch_experts_combine(golden):
    """Reference MOELayer.forward (src/UniGenUtils.py:74-191) + UniGenFlux.expert_forward (:925-967) vs the oracle's
    dispatch -> expert_forward -> combine on the same routing masks and expert weights."""
    g = golden["moe"]
    E, C = g["E"], g["C"]
    hidden, cond = g["hidden"], g["cond"]
    B, N, D = hidden.shape
    sd = {}
    for k, v in g["experts"].items():  # "e.br.j.weight" -> reference state-dict name
        sd[f"moe.moe_layer.experts.deepspeed_experts.{k}"] = v
    cfg = O.FluxConfig(num_attention_heads=1, attention_head_dim=D, condition_nums=0, expert_num_each_condition=E)
    assert cfg.expert_nums == E
    m = O.UniGenFluxOracle(cfg, sd)
    combine = g["combine"]
    dispatch = combine.bool()

    def disp(v):
        if v.dim() == 2:
            v = v[:, None, :].expand(-1, N, -1).reshape(-1, v.shape[-1])
        else:
            v = v.reshape(-1, v.shape[-1])
        return O.moe_dispatch(dispatch, v)[None]

    eh, ec = m.expert_forward(disp(hidden), disp(cond), disp(g["pooled"]), disp(g["cpooled"]))
    out_h = O.moe_combine(combine, eh.reshape(E, C, D), hidden)
    out_c = O.moe_combine(combine, ec.reshape(E, C, D), hidden)
    torch.testing.assert_close(out_h, g["out_hidden"], rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(out_c, g["out_cond"], rtol=1e-4, atol=1e-5)
    # dropped tokens are exactly zero
    dropped = combine.sum((1, 2)) == 0
    assert dropped.sum() == 2 and out_h.reshape(-1, D)[dropped].abs().max() == 0


def _affine_blocks(case):
    def mk_joint(i):
        a, b = 1.0 + 0.01 * (i + 1), 0.1 * (i + 1)
        return lambda h, c, t: (c * a + t[:, None] * 0.01, h * a + b + c.mean(1, keepdim=True) * 0.05 + t[:, None] * 0.02)

    def mk_single(i):
        a, b = 1.0 - 0.01 * (i + 1), -0.05 * (i + 1)
        return lambda x, t: x * a + b + t[:, None] * 0.03

    def lin(sd):
        return lambda x: torch.nn.functional.linear(x, sd["weight"], sd["bias"])

    nd, ns, dev = case["n_double"], case["n_single"], case["dev"]
    return ([mk_joint(i) for i in range(nd)], [mk_single(i) for i in range(ns)],
            [mk_joint(i + 50) for i in range(nd // dev)], [mk_single(i + 50) for i in range(ns // dev)],
            [lin(s) for s in case["adders_j"]], [lin(s) for s in case["adders_s"]])


def test_weave_matches_reference_base_forward(golden):
    """The reference's own base_forward/control_forward ran with affine stand-in blocks; the oracle's weave() must
    reproduce its outputs exactly-ish and its block call order exactly."""
    for case in golden["weave"]:
        bd, bs, cd, cs, ad, as_ = _affine_blocks(case)
        h, enc, _ = O.weave(case["h0"], case["e0"], case["temb"], lambda h_, e_: case["moe"], bd, bs, cd, cs, ad, as_,
                            0.7, case["method"])
        torch.testing.assert_close(h, case["out_hidden"], rtol=1e-5, atol=1e-5)
        torch.testing.assert_close(enc, case["out_ctx"], rtol=1e-5, atol=1e-5)
        # call order: base_d i, ctrl_d sched[i], ..., base_s i, ctrl_s sched[i]
        nd, ns, dev = case["n_double"], case["n_single"], case["dev"]
        want = []
        for i, j in enumerate(O.weave_schedule(nd, nd // dev)):
            want += [("base_d", i), ("ctrl_d", j + 50)]
        for i, j in enumerate(O.weave_schedule(ns, ns // dev)):
            want += [("base_s", i), ("ctrl_s", j + 50)]
        assert [tuple(c) for c in case["calls"]] == want


def test_moe_forward_wiring_consis_and_shared_experts(golden, monkeypatch):
    """The reference's own moe_forward (src/UniGenTransformer.py:969-1026) ran with affine stand-in joint blocks that fold
    their (hidden, encoder) ids into the result; the oracle's post_experts() must hand the same tensors, the same temb and the
    same id tables to the same blocks in the same order — with use_consis_module on / off and use_shared_expert on / off."""
    tag_of = {"consis_module.0": "consis0", "consis_module.1": "consis1", "shared_expert.0": "shared0", "shared_expert.1": "shared1"}
    for case in golden["moe_wiring"]:
        calls = []

        def stub_block(sd, prefix, H, h, c, temb, rope, trace=None, case=case, calls=calls):
            tag = tag_of[prefix]
            calls.append((tag, tuple(h.shape), tuple(c.shape)))
            n_enc = c.shape[1]
            enc_sig, hid_sig = rope[:n_enc].float().sum(-1), rope[n_enc:].float().sum(-1)  # encoder ids come first
            k = case["coef"][tag]
            out_enc = c * 0.9 + temb[:, None] * 0.01 + enc_sig[None, :, None] * 1e-3 + h.mean(1, keepdim=True) * 0.02 + k
            out_hid = h * 1.1 + temb[:, None] * 0.02 + hid_sig[None, :, None] * 2e-3 + c.mean(1, keepdim=True) * 0.05 - k
            return out_enc, out_hid

        monkeypatch.setattr(O, "flux_double_block", stub_block)
        monkeypatch.setattr(O, "flux_pos_embed", lambda ids, axes, theta: ids)  # the stand-ins consume the raw id tables
        D = case["hidden"].shape[-1]
        cfg = O.FluxConfig(num_attention_heads=1, attention_head_dim=D, use_shared_expert=case["shared"],
                           use_consis_module=case["consis"])
        m = O.UniGenFluxOracle(cfg, {})
        oh, oc = m.post_experts(case["hidden"], case["cond"], case["expert_hidden"], case["expert_cond"], case["enc"], case["temb"],
                                case["ctemb"], (case["img_ids"], case["txt_ids"], case["cond_ids"]))
        want_calls = [tuple(c) for c in case["calls"]]
        if case["consis"] and not case["shared"]:
            # the reference still RUNS the consis block there, but its result never reaches the return value (the tuple is only
            # rebuilt inside the shared-expert branch, :1024): the oracle skips the dead calls
            assert [c[0] for c in want_calls] == ["consis0", "consis0"] and calls == []
            torch.testing.assert_close(case["out_hidden"], case["expert_hidden"], rtol=0, atol=0)
        else:
            assert calls == want_calls
        torch.testing.assert_close(oh, case["out_hidden"], rtol=1e-6, atol=1e-6)
        torch.testing.assert_close(oc, case["out_cond"], rtol=1e-6, atol=1e-6)


def test_weave_schedule_integer_exact():
    assert O.weave_schedule(19, 9) == [0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8]
    assert O.weave_schedule(38, 19) == [i // 2 for i in range(38)]
    assert O.weave_schedule(2, 1) == [0, 0] and O.weave_schedule(4, 2) == [0, 0, 1, 1]


def test_enable_lora_semantics(golden):
    class FakeLora:
        def __init__(self, r, alpha, names):
            self.active_adapters = list(names)
            self.r = {n: r for n in names}
            self.lora_alpha = {n: alpha for n in names}
            self.scaling = {n: alpha / r for n in names}

        def set_scale(self, adapter, scale):
            if adapter in self.scaling:
                self.scaling[adapter] = scale * self.lora_alpha[adapter] / self.r[adapter]

    for case in golden["enable_lora"]:
        mods = [FakeLora(case["r"], case["alpha"], ["denoise", "depth", "canny"]), FakeLora(case["r"], case["alpha"], ["depth"]),
                object()]
        assert [dict(m.scaling) for m in mods[:2]] == case["before"]
        with O.enable_lora(mods, ["depth"]):
            assert [dict(m.scaling) for m in mods[:2]] == case["inside"]
        # NB: with alpha != r the reference's restore re-multiplies by alpha/r — replicated, not "fixed"
        assert [dict(m.scaling) for m in mods[:2]] == case["after"]
        assert [O.module_active_adapters(m) for m in mods] == case["active"]


def test_condition_ids_and_type_ids(golden):
    g = golden["condition"]
    assert g["condition_dict"] == O.CONDITION_DICT
    for case in g["cases"]:
        lh, lw = case["latent_hw"]
        ids, type_id = O.condition_ids(case["type"], lh * 8, lw * 8)  # latent = image / 8, token grid = latent / 2
        assert torch.equal(ids, case["ids"])
        assert torch.equal(type_id, case["type_id"])
