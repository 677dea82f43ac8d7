"""The reference's LoRA-switching hook over the native carriers (CPU: pure host logic). The oracle's `enable_lora` is pinned to
the real src/lora_switching_module.py by tests/test_oracle_golden.py; here the product-side hook must leave identical scaling
dicts on identical carriers at every phase (enter / exit / nesting / alpha != r)."""
import itertools

from oracle import unigen_oracle as O
from unigen_b200.lora_switching_module import LoraLayer, enable_lora, module_active_adapters


def _layers(alpha):
    ads = ["denoise", "depth", "canny"]
    mk = lambda: [LoraLayer(f"l{i}", ads, {a: 4 for a in ads}, dict(alpha)) for i in range(3)]  # noqa: E731
    return ads, mk(), mk()


def test_hook_matches_the_pinned_oracle_hook_phase_by_phase():
    for alpha in ({"denoise": 4, "depth": 4, "canny": 4}, {"denoise": 4, "depth": 2, "canny": 8}):
        ads, mine, theirs = _layers(alpha)
        for enabled in itertools.chain.from_iterable(itertools.combinations(ads, k) for k in range(4)):
            with enable_lora(mine + [object()], list(enabled)), O.enable_lora(theirs, list(enabled)):
                assert [m.scaling for m in mine] == [m.scaling for m in theirs]
                for m in mine:
                    assert all((m.scaling[a] == 0) == (a not in enabled) for a in ads)
                # nested context, as the predecessor's blocks nest them inside a user-level one
                with enable_lora(mine, ["depth"]), O.enable_lora(theirs, ["depth"]):
                    assert [m.scaling for m in mine] == [m.scaling for m in theirs]
                assert [m.scaling for m in mine] == [m.scaling for m in theirs]
            assert [m.scaling for m in mine] == [m.scaling for m in theirs]


def test_carrier_semantics():
    changed = []
    m = LoraLayer("x", ["a", "b"], {"a": 4, "b": 4}, {"a": 4.0, "b": 8.0}, on_change=changed.append)
    assert m.scaling == {"a": 1.0, "b": 2.0} and module_active_adapters(m) == ["a", "b"]
    m.set_scale("b", 0.5)           # peft: scale * lora_alpha / r
    assert m.scaling["b"] == 1.0 and changed == ["x"]
    m.set_scale("nope", 3.0)        # unknown adapter: ignored
    assert "nope" not in m.scaling
    m.set_adapter("a")
    assert m.active_adapters == ["a"] and m.effective_scale("b") == 0.0 and m.effective_scale("a") == 1.0
    m.scale_layer(0.5)
    assert m.scaling["a"] == 0.5 and m.scaling["b"] == 1.0  # only active adapters are scaled
    m.unscale_layer()
    assert m.scaling["a"] == 1.0
    assert module_active_adapters(object()) == []
