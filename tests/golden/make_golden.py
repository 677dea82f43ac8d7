"""Generate golden vectors by executing the REAL reference code from /root/reference (read-only, only in the build
container — it does not exist on the GPU box, hence the committed fixtures).

The reference imports diffusers / deepspeed / peft / ipdb / accelerate at module top; none is installable here
(SURVEY.md F3), so this script installs minimal stub modules that satisfy the imports and then drives the
reference-OWNED functions with real tensors:

  * `modulated_flatten`                       src/UniGenUtils.py:204-228   (both branches)
  * `MOELayer.forward` dispatch/combine       src/UniGenUtils.py:74-191    (gate = injected masks)
  * `UniGenFlux.expert_forward`               src/UniGenTransformer.py:925-967
  * `UniGenFlux.base_forward/control_forward` src/UniGenTransformer.py:1070-1180 with affine stand-in blocks
                                              (pins weave schedule, first-call substitution, overall_add / single_add)
  * `enable_lora` / `module_active_adapters`  src/lora_switching_module.py:4-39
  * `Condition._encode_image/encode`          src/condition.py:90-135      (ids, subject offset, type_id)

Third-party arithmetic is NOT exercised here (it does not exist on disk) and stays "parity unpinned".
Usage:  python tests/golden/make_golden.py     -> tests/golden/reference_golden.pt
"""
from __future__ import annotations

import sys
import types
from pathlib import Path

import torch
from torch import nn

REF = Path("/root/reference")
OUT = Path(__file__).resolve().parent / "reference_golden.pt"


def _stub(name: str, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    parent, _, child = name.rpartition(".")
    if parent:
        if parent not in sys.modules:
            _stub(parent)
        setattr(sys.modules[parent], child, m)
    return m


class _Anything:
    """Placeholder for third-party classes that are only imported / subclassed, never run."""

    def __init__(self, *a, **k):
        pass

    def __init_subclass__(cls, **k):
        pass


class _StubModule(nn.Module):
    def __init__(self, *a, **k):
        super().__init__()


def install_stubs():
    # --- deepspeed ---
    class StubMOELayer(nn.Module):  # deepspeed.moe.sharded_moe.MOELayer.__init__ signature
        def __init__(self, gate, experts, ep_group_name, ep_size, num_local_experts, use_tutel=False):
            super().__init__()
            self.gate, self.experts = gate, experts
            self.ep_group, self.ep_size, self.num_local_experts = None, ep_size, num_local_experts
            self.use_tutel, self.wall_clock_breakdown = use_tutel, False

    sharded = _stub("deepspeed.moe.sharded_moe", MOELayer=StubMOELayer, TopKGate=_StubModule, einsum=torch.einsum,
                    MOE_TIMER="moe", FIRST_ALLTOALL_TIMER="a2a1", SECOND_ALLTOALL_TIMER="a2a2")
    _stub("deepspeed.moe.mappings")
    _stub("deepspeed.moe.experts", Experts=_StubModule)
    _stub("deepspeed.moe", sharded_moe=sharded, mappings=sys.modules["deepspeed.moe.mappings"])
    groups = _stub("deepspeed.utils.groups", mpu=None)
    _stub("deepspeed.utils.bwc", bwc_tensor_model_parallel_world_size=lambda mpu=None: 1)
    _stub("deepspeed.utils", groups=groups)
    _stub("deepspeed")
    # --- diffusers (names only) ---
    names = {
        "diffusers.models.attention_processor": ["JointAttnProcessor2_0", "Attention", "FluxAttnProcessor2_0"],
        "diffusers.models.transformers.transformer_sd3": ["SD3SingleTransformerBlock", "SD3Transformer2DModel"],
        "diffusers.models.transformers.transformer_flux": ["FluxTransformer2DModel", "FluxTransformerBlock",
                                                           "FluxSingleTransformerBlock"],
        "diffusers.models.transformers.sana_transformer": ["SanaTransformer2DModel", "SanaTransformerBlock"],
        "diffusers.models.transformers": ["SD3Transformer2DModel", "FluxTransformer2DModel", "SanaTransformer2DModel"],
        "diffusers.models.attention": ["JointTransformerBlock", "FeedForward", "_chunked_feed_forward"],
        "diffusers.models.embeddings": ["FluxPosEmbed", "apply_rotary_emb", "PatchEmbed",
                                        "CombinedTimestepTextProjEmbeddings", "get_2d_sincos_pos_embed"],
        "diffusers.models.normalization": ["AdaLayerNormZero", "AdaLayerNormContinuous", "SD35AdaLayerNormZeroX",
                                           "RMSNorm"],
        "diffusers.models.modeling_outputs": ["Transformer2DModelOutput"],
        "diffusers.models": ["attention", "SD3Transformer2DModel", "FluxTransformer2DModel"],
        "diffusers.pipelines": ["FluxPipeline", "StableDiffusion3Pipeline"],
        "diffusers.utils": ["USE_PEFT_BACKEND", "logging", "scale_lora_layers", "unscale_lora_layers", "is_torch_version",
                            "deprecate"],
        "diffusers": ["FluxTransformer2DModel", "SD3Transformer2DModel", "FluxPipeline"],
    }
    for mod, ns in names.items():
        m = sys.modules.get(mod) or _stub(mod)
        for n in ns:
            if not hasattr(m, n):
                setattr(m, n, type(n, (_StubModule,), {}))
    sys.modules["diffusers.utils"].USE_PEFT_BACKEND = False
    sys.modules["diffusers.utils"].logging = types.SimpleNamespace(get_logger=lambda *_: types.SimpleNamespace(
        warning=lambda *a, **k: None, info=lambda *a, **k: None))
    sys.modules["diffusers.models"].attention = sys.modules["diffusers.models.attention"]

    # --- peft / misc ---
    class BaseTunerLayer:  # peft.tuners.tuners_utils.BaseTunerLayer surface used by enable_lora
        pass

    _stub("peft.tuners.tuners_utils", BaseTunerLayer=BaseTunerLayer)
    _stub("peft.tuners")
    _stub("peft")
    _stub("ipdb")
    _stub("cv2")
    _stub("tqdm", tqdm=lambda x, *a, **k: x)


class _ImportAnything(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        t = type(name, (_StubModule,), {})
        setattr(self, name, t)
        return t


def import_reference():
    """Import src.UniGenUtils / src.UniGenTransformer / src.lora_switching_module / src.condition from the reference,
    retrying with on-demand stubs for whatever third-party module or name is still missing."""
    install_stubs()
    sys.path.insert(0, str(REF))
    import importlib

    mods = {}
    for name in ("src.UniGenUtils", "src.lora_switching_module", "src.condition", "src.UniGenTransformer"):
        for _ in range(60):
            try:
                mods[name] = importlib.import_module(name)
                break
            except ModuleNotFoundError as e:
                missing = e.name
                sys.modules[missing] = _ImportAnything(missing)
                parent, _, child = missing.rpartition(".")
                if parent and parent in sys.modules:
                    setattr(sys.modules[parent], child, sys.modules[missing])
            except ImportError as e:  # cannot import name X from module Y
                msg = str(e)
                nm = msg.split("'")[1]
                modname = msg.split("'")[3] if msg.count("'") >= 4 else None
                if modname is None or modname not in sys.modules:
                    raise
                setattr(sys.modules[modname], nm, type(nm, (_StubModule,), {}))
        else:
            raise RuntimeError(f"could not import {name}")
    return mods


def main():
    mods = import_reference()
    U, T = mods["src.UniGenUtils"], mods["src.UniGenTransformer"]
    L, Cn = mods["src.lora_switching_module"], mods["src.condition"]
    g = torch.Generator().manual_seed(20240817)
    gold = {}

    # 1. modulated_flatten, both branches ------------------------------------------------------------------
    x = torch.randn(2, 5, 16, generator=g)
    w = torch.randn(24, 16, generator=g)
    s2 = torch.randn(2, 16, generator=g)
    s3 = torch.randn(2, 5, 16, generator=g)
    gold["modflat"] = dict(x=x, w=w, s2=s2, s3=s3, y2=U.modulated_flatten(x, w, s2).contiguous(),
                           y3=U.modulated_flatten(x, w, s3).contiguous())

    # 2. MOELayer.forward dispatch/combine + UniGenFlux.expert_forward --------------------------------------
    B, N, D, E, P, Tn = 2, 12, 16, 3, 8, 5
    S = B * N
    C = 9
    hidden, cond = torch.randn(B, N, D, generator=g), torch.randn(B, N, D, generator=g)
    enc = torch.randn(B, Tn, D, generator=g)
    temb, ctemb = torch.randn(B, D, generator=g), torch.randn(B, D, generator=g)
    pooled, cpooled = torch.randn(B, P, generator=g), torch.randn(B, P, generator=g)
    # an arbitrary but valid routing: token s -> expert s % E, slot by arrival order, drop two tokens
    combine = torch.zeros(S, E, C)
    probs = torch.rand(S, generator=g) * 0.5 + 0.3
    fill = [0] * E
    dropped = {4, 17}
    for s_ in range(S):
        e = s_ % E
        if s_ in dropped:
            continue
        combine[s_, e, fill[e]] = probs[s_]
        fill[e] += 1
    dispatch = combine.bool()
    experts = nn.ModuleList([
        nn.ModuleList([nn.ModuleList([nn.Linear(D, D), nn.Linear(P, D)]), nn.ModuleList([nn.Linear(D, D), nn.Linear(P, D)])])
        for _ in range(E)])
    with torch.no_grad():
        for p_ in experts.parameters():
            p_.copy_(torch.randn(p_.shape, generator=g) * 0.3)

    fake_self = types.SimpleNamespace(num_local_experts=E, use_modulate=False, use_rope=True,
                                      moe=types.SimpleNamespace(moe_layer=types.SimpleNamespace(
                                          experts=types.SimpleNamespace(deepspeed_experts=experts))))

    class ExpertsFn(nn.Module):
        def forward(self, **kw):
            return T.UniGenFlux.expert_forward(fake_self, **kw)

    class Gate(nn.Module):
        def forward(self, reshaped_input, used_token=None):
            return torch.tensor(0.25), combine, dispatch, torch.tensor(fill)

    layer = U.MOELayer(Gate(), ExpertsFn(), "ep_size_1", 1, E)
    with torch.no_grad():
        out_h, out_c = layer(choice_expert_input=hidden + cond, hidden_states=hidden, condition_hidden_states=cond,
                             used_token=None, encoder_hidden_states=enc, temb=temb, condition_temb=ctemb,
                             condition_pooled_projections=cpooled, pooled_projections=pooled)
    gold["moe"] = dict(hidden=hidden, cond=cond, enc=enc, temb=temb, ctemb=ctemb, pooled=pooled, cpooled=cpooled,
                       combine=combine, experts={k: v.detach().clone() for k, v in experts.state_dict().items()},
                       out_hidden=out_h.detach(), out_cond=out_c.detach(), E=E, C=C)

    # 3. weave: real base_forward + control_forward with affine stand-in blocks ------------------------------
    def run_weave(n_double, n_single, dev, method):
        Dm, Tm, Nm = 4, 3, 5
        gg = torch.Generator().manual_seed(7 + n_double * 100 + n_single)
        calls = []

        def mk_joint(tag, i):
            a, bcoef = 1.0 + 0.01 * (i + 1), 0.1 * (i + 1)

            def blk(hidden_states, encoder_hidden_states, temb, image_rotary_emb=None, joint_attention_kwargs=None):
                calls.append((tag, i))
                return (encoder_hidden_states * a + temb[:, None] * 0.01,
                        hidden_states * a + bcoef + encoder_hidden_states.mean(1, keepdim=True) * 0.05 + temb[:, None] * 0.02)
            return blk

        def mk_single(tag, i):
            a, bcoef = 1.0 - 0.01 * (i + 1), -0.05 * (i + 1)

            def blk(hidden_states, temb, image_rotary_emb=None, joint_attention_kwargs=None):
                calls.append((tag, i))
                return hidden_states * a + bcoef + temb[:, None] * 0.03
            return blk

        adders_j = [nn.Linear(Dm, Dm) for _ in range(n_double // dev)]
        adders_s = [nn.Linear(Dm, Dm) for _ in range(n_single // dev)]
        with torch.no_grad():
            for m_ in adders_j + adders_s:
                for p_ in m_.parameters():
                    p_.copy_(torch.randn(p_.shape, generator=gg) * 0.2)
        moe_out = dict(condition_hidden_states=torch.randn(1, Nm, Dm, generator=gg),
                       expert_hidden_states=torch.randn(1, Nm, Dm, generator=gg),
                       expert_condition_hidden_states=torch.randn(1, Nm, Dm, generator=gg),
                       control_encoder_hidden_states=torch.randn(1, Tm, Dm, generator=gg),
                       control_temb=torch.randn(1, Dm, generator=gg), condition_temb=torch.randn(1, Dm, generator=gg),
                       exp_count=torch.tensor([1, 2]), moe_loss=torch.tensor(0.5))
        fs = types.SimpleNamespace(
            transformer_blocks=[mk_joint("base_d", i) for i in range(n_double)],
            single_transformer_blocks=[mk_single("base_s", i) for i in range(n_single)],
            control_joint_trans_blocks=[mk_joint("ctrl_d", i + 50) for i in range(n_double // dev)],
            control_single_trans_blocks=[mk_single("ctrl_s", i + 50) for i in range(n_single // dev)],
            controlnet_add_joint_blocks=adders_j, controlnet_add_single_blocks=adders_s,
            single_block_control_method=method, use_rope=False)
        fs.preprocess_moe_forward = lambda *a, **k: dict(moe_out)
        fs.control_forward = lambda *a, **k: T.UniGenFlux.control_forward(fs, *a, **k)
        h0, e0 = torch.randn(1, Nm, Dm, generator=gg), torch.randn(1, Tm, Dm, generator=gg)
        temb_ = torch.randn(1, Dm, generator=gg)
        with torch.no_grad():
            res = T.UniGenFlux.base_forward(fs, hidden_states=h0, condition_hidden_states=torch.zeros(1, Nm, 2),
                                            encoder_hidden_states=e0, pooled_projections=None,
                                            condition_pooled_projections=None, timestep=None, conditioning_scale=0.7,
                                            temb=temb_, joint_attention_kwargs=None, image_rotary_emb=None, guidance=None,
                                            img_ids=None, prompt_ids=None, condition_ids=None)
        return dict(n_double=n_double, n_single=n_single, dev=dev, method=method, h0=h0, e0=e0, temb=temb_,
                    moe={k: v for k, v in moe_out.items()},
                    adders_j=[{k: v.detach().clone() for k, v in m_.state_dict().items()} for m_ in adders_j],
                    adders_s=[{k: v.detach().clone() for k, v in m_.state_dict().items()} for m_ in adders_s],
                    calls=calls, out_hidden=res["blocks_hidden_states"].detach(),
                    out_ctx=res["block_ctx_hidden_states"].detach())

    gold["weave"] = [run_weave(19, 38, 2, "overall_add"), run_weave(2, 4, 2, "overall_add"),
                     run_weave(5, 6, 2, "single_add")]

    # 3b. moe_forward wiring: the REAL UniGenFlux.moe_forward (src/UniGenTransformer.py:969-1026) with the MoE layer and the
    # joint blocks (consis_module / shared_expert) replaced by affine stand-ins that fold their ids into the result -----------
    def run_moe_wiring(consis, shared):
        Dm, Nm, Tm = 4, 5, 3
        gg = torch.Generator().manual_seed(31 + 2 * int(consis) + int(shared))
        calls = []
        coef = dict(consis0=0.3, consis1=0.7, shared0=-0.2, shared1=0.45)

        def mk_block(tag):
            def blk(hidden_states, encoder_hidden_states, temb=None, joint_attention_kwargs=None):
                jk = joint_attention_kwargs
                calls.append((tag, tuple(hidden_states.shape), tuple(encoder_hidden_states.shape)))
                hid_sig = jk["hd_ids"].float().sum(-1)
                enc_sig = jk["encoder_hd_ids"].float().sum(-1)
                out_enc = (encoder_hidden_states * 0.9 + temb[:, None] * 0.01 + enc_sig[None, :, None] * 1e-3
                           + hidden_states.mean(1, keepdim=True) * 0.02 + coef[tag])
                out_hid = (hidden_states * 1.1 + temb[:, None] * 0.02 + hid_sig[None, :, None] * 2e-3
                           + encoder_hidden_states.mean(1, keepdim=True) * 0.05 - coef[tag])
                return out_enc, out_hid
            return blk

        eh, ec = torch.randn(1, Nm, Dm, generator=gg), torch.randn(1, Nm, Dm, generator=gg)

        class Layer:
            l_aux = torch.tensor(0.125)
            exp_counts = torch.tensor([3, 2])

            def __call__(self, **kw):
                return eh, ec

        fs = types.SimpleNamespace(use_rope=True, pos_embed="pos_embed", use_consis_module=consis, use_shared_expert=shared,
                                   consis_module=[mk_block("consis0"), mk_block("consis1")],
                                   shared_expert=[mk_block("shared0"), mk_block("shared1")],
                                   moe=types.SimpleNamespace(moe_layer=Layer()))
        h, c = torch.randn(1, Nm, Dm, generator=gg), torch.randn(1, Nm, Dm, generator=gg)
        enc = torch.randn(1, Tm, Dm, generator=gg)
        temb, ctemb = torch.randn(1, Dm, generator=gg), torch.randn(1, Dm, generator=gg)
        img_ids = torch.randint(0, 9, (Nm, 3), generator=gg).float()
        txt_ids = torch.randint(0, 9, (Tm, 3), generator=gg).float()
        cond_ids = torch.randint(0, 9, (Nm, 3), generator=gg).float()
        with torch.no_grad():
            (oh, oc), l_aux, counts = T.UniGenFlux.moe_forward(
                fs, hidden_states=h, condition_hidden_states=c, encoder_hidden_states=enc, temb=temb, condition_temb=ctemb,
                condition_pooled_projections=None, pooled_projections=None,
                joint_attention_kwargs=dict(img_ids=img_ids, prompt_ids=txt_ids, condition_ids=cond_ids, rope_embed="rope"))
        return dict(consis=consis, shared=shared, hidden=h, cond=c, enc=enc, temb=temb, ctemb=ctemb, expert_hidden=eh,
                    expert_cond=ec, img_ids=img_ids, txt_ids=txt_ids, cond_ids=cond_ids, coef=coef, calls=calls,
                    out_hidden=oh.detach(), out_cond=oc.detach(), l_aux=l_aux, exp_counts=counts)

    gold["moe_wiring"] = [run_moe_wiring(True, True), run_moe_wiring(True, False), run_moe_wiring(False, True),
                          run_moe_wiring(False, False)]

    # 4. enable_lora ----------------------------------------------------------------------------------------
    Base = sys.modules["peft.tuners.tuners_utils"].BaseTunerLayer

    class FakeLora(Base):
        def __init__(self, r, alpha, names):
            self.active_adapters = list(names)
            self.r = {n: r for n in names}
            self.lora_alpha = {n: alpha for n in names}
            self.scaling = {n: alpha / r for n in names}

        def set_scale(self, adapter, scale):  # peft 0.15 LoraLayer.set_scale
            if adapter not in self.scaling:
                return
            self.scaling[adapter] = scale * self.lora_alpha[adapter] / self.r[adapter]

    lora_cases = []
    for r, alpha in ((4, 4), (4, 8)):
        mods_ = [FakeLora(r, alpha, ["denoise", "depth", "canny"]), FakeLora(r, alpha, ["depth"]), object()]
        before = [dict(m.scaling) for m in mods_[:2]]
        with L.enable_lora(mods_, ["depth"]):
            inside = [dict(m.scaling) for m in mods_[:2]]
        after = [dict(m.scaling) for m in mods_[:2]]
        lora_cases.append(dict(r=r, alpha=alpha, before=before, inside=inside, after=after,
                               active=[L.module_active_adapters(m) for m in mods_]))
    gold["enable_lora"] = lora_cases

    # 5. Condition ids / type ids ---------------------------------------------------------------------------
    class FakeImg:
        def to(self, *a, **k):
            return self

    def pack_latents(latents, b, c, h, w):  # diffusers FluxPipeline._pack_latents (third-party, restated)
        latents = latents.view(b, c, h // 2, 2, w // 2, 2).permute(0, 2, 4, 1, 3, 5)
        return latents.reshape(b, (h // 2) * (w // 2), c * 4)

    def prep_ids(b, h, w, device, dtype):  # diffusers FluxPipeline._prepare_latent_image_ids (third-party, restated)
        ids = torch.zeros(h, w, 3)
        ids[..., 1] = ids[..., 1] + torch.arange(h)[:, None]
        ids[..., 2] = ids[..., 2] + torch.arange(w)[None, :]
        return ids.reshape(h * w, 3)

    cond_cases = []
    for ctype, (lh, lw) in (("canny", (8, 8)), ("depth", (6, 10)), ("subject", (6, 10)), ("subject", (8, 4))):
        lat = torch.randn(1, 16, lh, lw, generator=g)
        pipe = types.SimpleNamespace(
            image_processor=types.SimpleNamespace(preprocess=lambda im: FakeImg()),
            vae=types.SimpleNamespace(encode=lambda im, lat=lat: types.SimpleNamespace(
                latent_dist=types.SimpleNamespace(sample=lambda: lat)),
                config=types.SimpleNamespace(shift_factor=0.1159, scaling_factor=0.3611)),
            _pack_latents=pack_latents, _prepare_latent_image_ids=prep_ids, device="cpu", dtype=torch.float32)
        c = Cn.Condition(ctype, raw_img=types.SimpleNamespace(convert=lambda mode: "img"), no_process=True)
        tokens, ids, type_id = c.encode(pipe)
        cond_cases.append(dict(type=ctype, latent_hw=(lh, lw), ids=ids, type_id=type_id, tokens=tokens, latents=lat))
    gold["condition"] = dict(cases=cond_cases, condition_dict=dict(Cn.condition_dict))

    torch.save(gold, OUT)
    print("wrote", OUT, {k: (len(v) if isinstance(v, (list, dict)) else type(v)) for k, v in gold.items()})


if __name__ == "__main__":
    main()
