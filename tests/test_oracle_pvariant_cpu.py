"""CPU checks of the P-variant oracle pieces: the `enable_lora` hook semantics (pinned against the real reference hook in
test_oracle_golden.py) agree with the data form the native path uses (adapter group per row segment)."""
import torch

from oracle import unigen_oracle as O


class PeftLikeLinear:
    """peft 0.15 lora.Linear surface: active_adapters, scaling, set_scale, forward = base + sum B(A x) * scaling."""

    def __init__(self, weight, bias, adapters, r=4, alpha=4, seed=0):
        g = torch.Generator().manual_seed(seed)
        self.weight, self.bias = weight, bias
        self.active_adapters = list(adapters)
        self.r = {a: r for a in adapters}
        self.lora_alpha = {a: alpha for a in adapters}
        self.scaling = {a: alpha / r for a in adapters}
        self.A = {a: torch.randn(r, weight.shape[1], generator=g) for a in adapters}
        self.B = {a: torch.randn(weight.shape[0], r, generator=g) for a in adapters}

    def set_scale(self, adapter, scale):
        if adapter in self.scaling:
            self.scaling[adapter] = scale * self.lora_alpha[adapter] / self.r[adapter]

    def __call__(self, x):
        return O.lora_linear(x, self.weight, self.bias, {a: (self.A[a], self.B[a], self.scaling[a]) for a in self.active_adapters},
                             self.active_adapters)


def test_enable_lora_context_equals_segment_group_form():
    g = torch.Generator().manual_seed(1)
    W, b = torch.randn(24, 16, generator=g), torch.randn(24, generator=g)
    adapters, cond_types = ["denoise", "depth", "canny"], ["depth", "canny"]
    lin = PeftLikeLinear(W, b, adapters)
    sd = {"l.weight": W, "l.bias": b}
    for a in adapters:
        sd[f"l.lora_A.{a}.weight"], sd[f"l.lora_B.{a}.weight"] = lin.A[a], lin.B[a]
    pv = O.PVariantOracle(O.FluxConfig.tiny(), sd, adapters)
    x = torch.randn(3, 5, 16, generator=g)
    den = [a for a in adapters if a not in cond_types]
    for active in (den, ["depth"], ["canny"], []):
        with O.enable_lora([lin], active):
            want = lin(x)
        torch.testing.assert_close(pv.lin("l", x, active), want, rtol=1e-5, atol=1e-5)
    assert lin.scaling == {a: 1.0 for a in adapters}  # restored (alpha == r)


def test_segment_mask_rules():
    vis = O.pvariant_visibility(2)
    m = O.segment_mask([0, 2, 5, 7, 9], vis)
    assert m[:5].all()                                   # txt, img rows see every key
    assert m[5:7, :5].all() and m[5:7, 5:7].all() and not m[5:7, 7:].any()   # c_1 sees txt, img, c_1 only
    assert m[7:, :5].all() and m[7:, 7:].all() and not m[7:, 5:7].any()
    ms = O.segment_mask([0, 2, 5, 7, 9], O.pvariant_visibility(2, strict=True))
    assert not ms[5:7, :5].any() and ms[5:7, 5:7].all()  # north-star wording: condition tokens see only themselves
    assert ms[:5].all()
