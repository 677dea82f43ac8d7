"""Product-side host logic against the golden vectors generated from the REAL reference functions
(tests/golden/make_golden.py imports src/condition.py and src/lora_switching_module.py from /root/reference):
`unigen_b200.condition` ids / type ids and `unigen_b200.lora_switching_module` are bit-exact with them."""
import torch

from unigen_b200 import condition as PC
from unigen_b200.lora_switching_module import LoraLayer, enable_lora, module_active_adapters


def test_condition_ids_and_type_ids_bit_exact(golden):
    g = golden["condition"]
    assert g["condition_dict"] == PC.condition_dict
    assert len(g["cases"]) >= 3
    for case in g["cases"]:
        lh, lw = case["latent_hw"]
        ids, type_id = PC.condition_ids(case["type"], lh * 8, lw * 8)
        assert torch.equal(ids, case["ids"]) and torch.equal(type_id, case["type_id"])
        assert ids.dtype == case["ids"].dtype
        # bf16 (the pipeline dtype) holds the same integers exactly
        assert torch.equal(PC.condition_ids(case["type"], lh * 8, lw * 8, dtype=torch.bfloat16)[0].float(), case["ids"])
        # the class form: tokens + ids + type_id as Condition.encode returns them
        tokens = torch.zeros(1, ids.shape[0], 64)
        t, i, ty = PC.Condition(case["type"], tokens, height=lh * 8, width=lw * 8).encode()
        assert t is tokens and torch.equal(i, case["ids"]) and torch.equal(ty, case["type_id"])
        assert PC.Condition(case["type"], tokens, condition_ids=case["ids"]).encode()[1] is case["ids"]


def test_latent_image_ids_match_the_oracle_restatement():
    from oracle import unigen_oracle as O
    for h, w in ((16, 16), (64, 64), (20, 12), (1, 7)):
        assert torch.equal(PC.prepare_latent_image_ids(h, w), O.prepare_latent_image_ids(h, w))


def test_enable_lora_golden_cases(golden):
    """Same replay as tests/test_oracle_golden.py, on the product hook + native carriers."""
    for case in golden["enable_lora"]:
        r, alpha = case["r"], case["alpha"]

        def carrier(adapters):
            return LoraLayer("m", adapters, {a: r for a in adapters}, {a: alpha for a in adapters})

        mods = [carrier(["denoise", "depth", "canny"]), carrier(["depth"]), object()]
        assert [dict(m.scaling) for m in mods[:2]] == case["before"]
        with enable_lora(mods, ["depth"]):
            assert [dict(m.scaling) for m in mods[:2]] == case["inside"]
        assert [dict(m.scaling) for m in mods[:2]] == case["after"]
        assert [module_active_adapters(m) for m in mods] == case["active"]
