"""LoRA-switching hooks of the reference, over the B200-native LoRA carriers.

Reference: src/lora_switching_module.py:4-39 — `module_active_adapters(module)` and the `enable_lora(lora_modules,
enable_adapters)` context manager, which the predecessor's blocks wrap around every switched linear
(UniCombineTransformerBlock.pyc L22, 81, 121, 130, 222, 229, 251, 261, 283, 287). There the modules are PEFT
`BaseTunerLayer`s; here they are `LoraLayer` carriers, one per LoRA-wrapped linear of `UniCombineFlux`
(`model.lora_layers[name]`), exposing the same three members the hook touches:

    .active_adapters : list of adapter names applied by the linear's forward
    .scaling         : dict adapter -> scale (peft: lora_alpha / r)
    .set_scale(a, s) : peft 0.15 LoraLayer.set_scale — stores  s * lora_alpha[a] / r[a]

The native forward never enters a context per segment — the per-segment switch is DATA (adapter-group table consumed by the
GEMM, pvariant.py). What the hook changes is the scale table those GEMMs read: a carrier whose `.scaling` /
`.active_adapters` changed marks its linear dirty and `UniCombineFlux` rebuilds the pre-scaled low-rank operands (in place,
so captured CUDA graphs stay valid) before the next forward. The restore quirk of the reference is kept as is: `__exit__`
calls `set_scale(adapter, saved_scaling)`, which multiplies by lora_alpha / r AGAIN, so the round trip is exact only when
lora_alpha == r (the UniCombine convention, train.py:137) — SURVEY.md §8 A14 "replicate, don't fix"."""
from __future__ import annotations

from typing import Any, Callable, Dict, Iterable, List, Optional, Sequence


class LoraLayer:
    """Native stand-in for a PEFT LoRA `Linear` (peft 0.15 tuners/lora/layer.py): adapter bookkeeping only — the low-rank
    weights live in the owning model's fused operand stacks."""

    def __init__(self, name: str, adapters: Sequence[str], r: Dict[str, int], lora_alpha: Dict[str, float],
                 on_change: Optional[Callable[[str], None]] = None):
        self.name = name
        self.r = dict(r)
        self.lora_alpha = dict(lora_alpha)
        self._on_change = on_change
        self._active: List[str] = list(adapters)
        self.scaling: Dict[str, float] = {a: self.lora_alpha[a] / self.r[a] for a in adapters}

    # peft BaseTunerLayer.active_adapters is a property over `_active_adapter`; assignment goes through set_adapter
    @property
    def active_adapters(self) -> List[str]:
        return list(self._active)

    def set_adapter(self, adapter_names) -> None:
        self._active = [adapter_names] if isinstance(adapter_names, str) else list(adapter_names)
        self._changed()

    def set_scale(self, adapter: str, scale: float) -> None:
        """peft LoraLayer.set_scale: ignored for unknown adapters, otherwise scaling = scale * lora_alpha / r."""
        if adapter not in self.scaling:
            return
        self.scaling[adapter] = scale * self.lora_alpha[adapter] / self.r[adapter]
        self._changed()

    def scale_layer(self, scale: float) -> None:
        """peft LoraLayer.scale_layer (what `joint_attention_kwargs['scale']` reaches through scale_lora_layers)."""
        if scale == 1:
            return
        for a in self._active:
            if a in self.scaling:
                self.scaling[a] *= scale
        self._changed()

    def unscale_layer(self, scale: Optional[float] = None) -> None:
        for a in self._active:
            if a not in self.scaling:
                continue
            if scale is None:
                self.scaling[a] = self.lora_alpha[a] / self.r[a]
            else:
                self.scaling[a] /= scale
        self._changed()

    def effective_scale(self, adapter: str) -> float:
        """Scale the linear's forward applies to `adapter`: 0 unless it is active."""
        return float(self.scaling.get(adapter, 0.0)) if adapter in self._active else 0.0

    def _changed(self) -> None:
        if self._on_change is not None:
            self._on_change(self.name)


def module_active_adapters(module: Any) -> List[str]:
    """src/lora_switching_module.py:4-9: the active adapters that carry a scale."""
    if not hasattr(module, "active_adapters"):
        return []
    known = module.scaling.keys()
    return [a for a in module.active_adapters if a in known]


class enable_lora:
    """src/lora_switching_module.py:11-39. Inside the context only `enable_adapters` keep their scale on `lora_modules`."""

    def __init__(self, lora_modules: Iterable[Any], enable_adapters: Sequence[str]) -> None:
        self.lora_modules = [m for m in lora_modules if isinstance(m, LoraLayer) or
                             (hasattr(m, "set_scale") and hasattr(m, "scaling") and hasattr(m, "active_adapters"))]
        self.active_adapter_scales = [{a: m.scaling[a] for a in module_active_adapters(m)} for m in self.lora_modules]
        self.enable_adapters = enable_adapters

    def __enter__(self) -> None:
        for m in self.lora_modules:
            for a in module_active_adapters(m):
                if a not in self.enable_adapters:
                    m.set_scale(a, 0)

    def __exit__(self, exc_type, exc_val, exc_tb) -> None:
        for m, saved in zip(self.lora_modules, self.active_adapter_scales):
            for a in module_active_adapters(m):
                m.set_scale(a, saved[a])  # re-multiplied by lora_alpha / r inside set_scale (reference behaviour)
