"""ORACLE — test infrastructure, NOT product code.

CPU-runnable, pure-torch restatement of the reference hot path: one `UniGenFlux.forward` (= one denoise step) of
gavin-gqzhang/UniGen, canonical configuration CANON-FLUX of SURVEY.md §A.1.  Only tests/, `__graft_entry__.smoke()`
and `bench.py`'s cpu_baseline / `--impl reference` legs may import this file; the product path
(`unigen_b200/`) never does and fails loudly when libunigen_b200.so is missing.

Parity pinning status
  * reference-owned arithmetic — `modulated_flatten`, the expert math, the GShard dispatch/combine algebra of
    `MOELayer`, `enable_lora`, `Condition` ids — is PINNED: tests/golden/make_golden.py imports the real functions
    from /root/reference (with stub modules for the absent third-party imports) and commits input/output vectors
    that tests/test_oracle_golden.py replays against this file.
  * third-party arithmetic (diffusers 0.32.2 Flux blocks / AdaLN / RoPE / embeddings, deepspeed 0.16.5 top1gating,
    peft 0.15 LoRA) is "parity unpinned": the packages are not installable here (no network, not in the wheelhouse)
    and the reference has no tests or golden vectors; the restatement follows the published algorithms of those
    pinned versions as written down in SURVEY.md Appendix A, and the op ORDER is cross-checked against the
    reference's own predecessor bytecode (src/__pycache__/UniCombineTransformerBlock.cpython-312.pyc).

Every function cites the reference file:line (relative to /root/reference) or the SURVEY appendix it follows.
State-dict keys are the reference's (diffusers Flux names + `control_*`, `controlnet_add_*`, `moe.*`, `shared_expert.*`,
src/UniGenTransformer.py:728-773,861,891).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# ------------------------------------------------------------------------------------------------------------------
# configuration
# ------------------------------------------------------------------------------------------------------------------
@dataclass
class FluxConfig:
    """diffusers FluxTransformer2DModel config fields the path reads + UniGen control_params."""
    num_layers: int = 19
    num_single_layers: int = 38
    attention_head_dim: int = 128
    num_attention_heads: int = 24
    in_channels: int = 64
    joint_attention_dim: int = 4096
    pooled_projection_dim: int = 768
    guidance_embeds: bool = False
    axes_dims_rope: Tuple[int, int, int] = (16, 56, 56)
    theta: float = 10000.0
    # control_params (config/unigen.yaml:3-11 + use_rope: True)
    condition_nums: int = 1
    expert_num_each_condition: int = 3
    single_control_dev: int = 2
    single_block_control_method: str = "overall_add"
    use_pooled_prompt_embeds: bool = True
    use_shared_expert: bool = True
    use_consis_module: bool = False  # src/UniGenTransformer.py:893-923 (V2: two joint blocks, the first one used twice)

    @property
    def inner_dim(self) -> int:
        return self.num_attention_heads * self.attention_head_dim

    @property
    def expert_nums(self) -> int:  # src/UniGenTransformer.py:807
        return (self.condition_nums + 1) * self.expert_num_each_condition

    @property
    def cn_joint_layers(self) -> int:  # src/UniGenTransformer.py:744
        return self.num_layers // self.single_control_dev

    @property
    def cn_single_layers(self) -> int:
        return self.num_single_layers // self.single_control_dev

    @staticmethod
    def tiny() -> "FluxConfig":
        """cfg1 of BASELINE.json / SURVEY.md §8(d): 2 double + 4 single blocks, hidden 384, 6 heads of 64."""
        return FluxConfig(num_layers=2, num_single_layers=4, attention_head_dim=64, num_attention_heads=6,
                          axes_dims_rope=(8, 28, 28))

    @staticmethod
    def flux() -> "FluxConfig":
        """FLUX.1-schnell architecture (cfg2-cfg4)."""
        return FluxConfig()


# ------------------------------------------------------------------------------------------------------------------
# third-party pieces restated
# ------------------------------------------------------------------------------------------------------------------
def linear(sd: Dict[str, Tensor], prefix: str, x: Tensor) -> Tensor:
    return F.linear(x, sd[prefix + ".weight"], sd.get(prefix + ".bias"))


def timesteps_proj(t: Tensor, dim: int = 256) -> Tensor:
    """diffusers Timesteps(256, flip_sin_to_cos=True, downscale_freq_shift=0)."""
    half = dim // 2
    exponent = -math.log(10000.0) * torch.arange(half, dtype=torch.float32, device=t.device) / half
    emb = t[:, None].float() * torch.exp(exponent)[None, :]
    emb = torch.cat([torch.sin(emb), torch.cos(emb)], dim=-1)
    return torch.cat([emb[:, half:], emb[:, :half]], dim=-1)  # flip_sin_to_cos -> [cos | sin]


def combined_timestep_text_embed(sd, prefix: str, timestep: Tensor, pooled: Tensor,
                                 guidance: Optional[Tensor] = None) -> Tensor:
    """CombinedTimestep(Guidance)TextProjEmbeddings; called at src/UniGenTransformer.py:1222,1048-1049."""
    t = timesteps_proj(timestep).to(pooled.dtype)
    t = linear(sd, prefix + ".timestep_embedder.linear_2", F.silu(linear(sd, prefix + ".timestep_embedder.linear_1", t)))
    if guidance is not None:
        g = timesteps_proj(guidance).to(pooled.dtype)
        t = t + linear(sd, prefix + ".guidance_embedder.linear_2",
                       F.silu(linear(sd, prefix + ".guidance_embedder.linear_1", g)))
    p = linear(sd, prefix + ".text_embedder.linear_2", F.silu(linear(sd, prefix + ".text_embedder.linear_1", pooled)))
    return t + p


def flux_pos_embed(ids: Tensor, axes_dim: Tuple[int, ...], theta: float = 10000.0) -> Tuple[Tensor, Tensor]:
    """FluxPosEmbed / get_1d_rotary_pos_embed(repeat_interleave_real=True, use_real=True)."""
    cos_out, sin_out = [], []
    pos = ids.double()
    for i, d in enumerate(axes_dim):
        freqs = 1.0 / (theta ** (torch.arange(0, d, 2, dtype=torch.float64, device=ids.device)[: d // 2] / d))
        ang = torch.outer(pos[:, i], freqs)
        cos_out.append(ang.cos().repeat_interleave(2, dim=1).float())
        sin_out.append(ang.sin().repeat_interleave(2, dim=1).float())
    return torch.cat(cos_out, dim=-1), torch.cat(sin_out, dim=-1)


def apply_rotary_emb(x: Tensor, rope: Tuple[Tensor, Tensor]) -> Tensor:
    """diffusers apply_rotary_emb(use_real=True, use_real_unbind_dim=-1); x: (B,H,S,dh)."""
    cos, sin = rope
    cos, sin = cos[None, None], sin[None, None]
    x_real, x_imag = x.reshape(*x.shape[:-1], -1, 2).unbind(-1)
    x_rot = torch.stack([-x_imag, x_real], dim=-1).flatten(3)
    return (x.float() * cos + x_rot.float() * sin).to(x.dtype)


def rms_norm(x: Tensor, weight: Tensor, eps: float = 1e-6) -> Tensor:
    """diffusers RMSNorm."""
    in_dtype = x.dtype
    var = x.float().pow(2).mean(-1, keepdim=True)
    x = x * torch.rsqrt(var + eps)
    if weight.dtype in (torch.float16, torch.bfloat16):
        x = x.to(weight.dtype)
    return (x * weight).to(in_dtype) if weight.dtype == in_dtype else x * weight


def layer_norm(x: Tensor, eps: float = 1e-6) -> Tensor:
    return F.layer_norm(x, (x.shape[-1],), None, None, eps)


def gelu_tanh(x: Tensor) -> Tensor:
    return F.gelu(x, approximate="tanh")


def feed_forward(sd, prefix: str, x: Tensor) -> Tensor:
    """FeedForward(dim, dim, activation_fn='gelu-approximate')."""
    return linear(sd, prefix + ".net.2", gelu_tanh(linear(sd, prefix + ".net.0.proj", x)))


def _heads(x: Tensor, H: int) -> Tensor:
    B, S, D = x.shape
    return x.view(B, S, H, D // H).transpose(1, 2)


def sdpa(q: Tensor, k: Tensor, v: Tensor, mask: Optional[Tensor] = None) -> Tensor:
    """F.scaled_dot_product_attention(dropout_p=0, is_causal=False); written out so that fp32 CPU runs are
    deterministic across torch versions. q,k,v: (B,H,S,dh); mask: bool (Sq,Sk), True = attend."""
    scale = 1.0 / math.sqrt(q.shape[-1])
    s = torch.matmul(q.float(), k.float().transpose(-1, -2)) * scale
    if mask is not None:
        s = s.masked_fill(~mask, float("-inf"))
    p = torch.softmax(s, dim=-1)
    return torch.matmul(p, v.float()).to(q.dtype)


def flux_double_block(sd, prefix: str, H: int, h: Tensor, c: Tensor, temb: Tensor,
                      rope: Optional[Tuple[Tensor, Tensor]], trace: Optional[dict] = None) -> Tuple[Tensor, Tensor]:
    """diffusers FluxTransformerBlock.forward + FluxAttnProcessor2_0 (op order confirmed by
    UniCombineTransformerBlock.pyc L20-136, L152-236). Returns (encoder_hidden_states, hidden_states)."""
    e = linear(sd, prefix + ".norm1.linear", F.silu(temb))
    shift_msa, scale_msa, gate_msa, shift_mlp, scale_mlp, gate_mlp = e.chunk(6, dim=1)
    nh = layer_norm(h) * (1 + scale_msa[:, None]) + shift_msa[:, None]
    e = linear(sd, prefix + ".norm1_context.linear", F.silu(temb))
    c_shift_msa, c_scale_msa, c_gate_msa, c_shift_mlp, c_scale_mlp, c_gate_mlp = e.chunk(6, dim=1)
    nc = layer_norm(c) * (1 + c_scale_msa[:, None]) + c_shift_msa[:, None]

    a = prefix + ".attn"
    q, k, v = (_heads(linear(sd, f"{a}.to_{n}", nh), H) for n in "qkv")
    q, k = rms_norm(q, sd[a + ".norm_q.weight"]), rms_norm(k, sd[a + ".norm_k.weight"])
    cq, ck, cv = (_heads(linear(sd, f"{a}.add_{n}_proj", nc), H) for n in "qkv")
    cq, ck = rms_norm(cq, sd[a + ".norm_added_q.weight"]), rms_norm(ck, sd[a + ".norm_added_k.weight"])
    Q, K, V = torch.cat([cq, q], 2), torch.cat([ck, k], 2), torch.cat([cv, v], 2)  # TEXT FIRST
    if rope is not None:
        Q, K = apply_rotary_emb(Q, rope), apply_rotary_emb(K, rope)
    O = sdpa(Q, K, V)
    B, _, S, dh = O.shape
    O = O.transpose(1, 2).reshape(B, S, H * dh).to(q.dtype)
    T = c.shape[1]
    ca, at = O[:, :T], O[:, T:]
    at = linear(sd, a + ".to_out.0", at)
    ca = linear(sd, a + ".to_add_out", ca)
    if trace is not None:
        trace.update(norm_hidden=nh, norm_context=nc, attn_out=O)

    h = h + gate_msa[:, None] * at
    h = h + gate_mlp[:, None] * feed_forward(sd, prefix + ".ff", layer_norm(h) * (1 + scale_mlp[:, None]) + shift_mlp[:, None])
    c = c + c_gate_msa[:, None] * ca
    c = c + c_gate_mlp[:, None] * feed_forward(sd, prefix + ".ff_context",
                                               layer_norm(c) * (1 + c_scale_mlp[:, None]) + c_shift_mlp[:, None])
    if c.dtype == torch.float16:
        c = c.clip(-65504, 65504)
    return c, h


def flux_single_block(sd, prefix: str, H: int, x: Tensor, temb: Tensor,
                      rope: Optional[Tuple[Tensor, Tensor]]) -> Tensor:
    """diffusers FluxSingleTransformerBlock.forward (pyc L250-295)."""
    residual = x
    e = linear(sd, prefix + ".norm.linear", F.silu(temb))
    shift, scale, gate = e.chunk(3, dim=1)
    nx = layer_norm(x) * (1 + scale[:, None]) + shift[:, None]
    m = gelu_tanh(linear(sd, prefix + ".proj_mlp", nx))
    a = prefix + ".attn"
    q, k, v = (_heads(linear(sd, f"{a}.to_{n}", nx), H) for n in "qkv")
    q, k = rms_norm(q, sd[a + ".norm_q.weight"]), rms_norm(k, sd[a + ".norm_k.weight"])
    if rope is not None:
        q, k = apply_rotary_emb(q, rope), apply_rotary_emb(k, rope)
    O = sdpa(q, k, v)
    B, _, S, dh = O.shape
    O = O.transpose(1, 2).reshape(B, S, H * dh).to(x.dtype)
    x = residual + gate[:, None] * linear(sd, prefix + ".proj_out", torch.cat([O, m], dim=2))
    if x.dtype == torch.float16:
        x = x.clip(-65504, 65504)
    return x


def ada_layer_norm_continuous(sd, prefix: str, x: Tensor, temb: Tensor) -> Tensor:
    """AdaLayerNormContinuous: scale FIRST, then shift."""
    e = linear(sd, prefix + ".linear", F.silu(temb).to(x.dtype))
    scale, shift = e.chunk(2, dim=1)
    return layer_norm(x) * (1 + scale)[:, None] + shift[:, None]


# ------------------------------------------------------------------------------------------------------------------
# DeepSpeed 0.16.5 top1gating as used here — k=1, capacity_factor=1, min_capacity=4, use_rts=True
# ------------------------------------------------------------------------------------------------------------------
def moe_capacity(tokens: int, experts: int, capacity_factor: float = 1.0, min_capacity: int = 4) -> int:
    return max(int(math.ceil(tokens / experts * capacity_factor)), min_capacity)


def top1gating(logits: Tensor, capacity: int, rts_uniform: Tensor):
    """Returns (l_aux, combine_weights (S,E,C) fp32, dispatch_mask bool, exp_counts int64[E]) plus the sparse view
    (expert_idx, slot (-1 = dropped), prob) the CUDA path produces."""
    S, E = logits.shape
    gates = F.softmax(logits, dim=1)
    idx = torch.argmax(gates, dim=1)
    mask1 = F.one_hot(idx, num_classes=E)
    exp_counts = mask1.sum(dim=0).detach()
    me = gates.mean(dim=0)
    ce = mask1.float().mean(dim=0)
    l_aux = torch.sum(me * ce) * E
    mask1_rand = mask1 * rts_uniform
    top_idx = torch.topk(mask1_rand, k=capacity, dim=0)[1]
    new_mask1 = mask1 * torch.zeros_like(mask1).scatter_(0, top_idx, 1)
    mask1 = new_mask1
    locations1 = torch.cumsum(mask1, dim=0) - 1
    locations1_s = torch.sum(locations1 * mask1, dim=1)
    mask1_float = mask1.float()
    gates = gates * mask1_float
    locations1_sc = F.one_hot(locations1_s, num_classes=capacity).float()
    combine_weights = torch.einsum("se,sc->sec", gates, locations1_sc)
    dispatch_mask = combine_weights.bool()
    kept = mask1.sum(dim=1) > 0
    slot = torch.where(kept, locations1_s, torch.full_like(locations1_s, -1))
    prob = F.softmax(logits, dim=1).gather(1, idx[:, None])[:, 0]
    return l_aux, combine_weights, dispatch_mask, exp_counts, (idx, slot, prob)


# ------------------------------------------------------------------------------------------------------------------
# reference-owned pieces restated
# ------------------------------------------------------------------------------------------------------------------
def modulated_flatten(x: Tensor, w: Tensor, s: Tensor) -> Tensor:
    """src/UniGenUtils.py:204-228. y[b,n,:] = (W * s[b(,n)]) x[b,n,:]  ==  W (s ⊙ x)."""
    if s.dim() == 2:
        s = s[:, None, :]
    return torch.matmul(x * s, w.t())


def moe_dispatch(dispatch_mask: Tensor, value: Tensor) -> Tensor:
    """src/UniGenUtils.py:136-159: einsum('sec,sm->ecm') with the one-hot mask cast to the value dtype."""
    return torch.einsum("sec,sm->ecm", dispatch_mask.type_as(value), value)


def moe_combine(combine_weights: Tensor, expert_output: Tensor, like: Tensor) -> Tensor:
    """src/UniGenUtils.py:161-191: einsum('sec,ecm->sm') then reshape to the routed input's shape."""
    E, C, M = expert_output.shape
    return torch.einsum("sec,ecm->sm", combine_weights.type_as(like), expert_output).reshape(like.shape)


def prepare_latent_image_ids(height: int, width: int) -> Tensor:
    """FluxPipeline._prepare_latent_image_ids(batch, h//2, w//2) -> (h*w, 3) rows (0,row,col);
    call sites src/condition.py:101-108, src/UniGenPipeline.py:640-647."""
    ids = torch.zeros(height, width, 3)
    ids[..., 1] = ids[..., 1] + torch.arange(height)[:, None]
    ids[..., 2] = ids[..., 2] + torch.arange(width)[None, :]
    return ids.reshape(height * width, 3)


CONDITION_DICT = {"depth": 0, "canny": 1, "subject": 4, "coloring": 6, "deblurring": 7, "fill": 9}  # src/condition.py:12-19


def condition_ids(condition_type: str, height: int, width: int) -> Tuple[Tensor, Tensor]:
    """Token-index part of Condition._encode_image / encode (src/condition.py:101-111,134): latent grid of
    (height//16, width//16) packed tokens; `subject` shifts column ids by the grid width; type_id per token."""
    h2, w2 = height // 16, width // 16
    ids = prepare_latent_image_ids(h2, w2)
    if condition_type == "subject":
        ids[:, 2] += h2  # `cond_ids[:, 2] += cond_img.shape[2] // 2` — the latent HEIGHT // 2 (src/condition.py:109-110)
    type_id = torch.ones_like(ids[:, :1]) * CONDITION_DICT[condition_type]
    return ids, type_id


def weave_schedule(n_base: int, n_ctrl: int) -> List[int]:
    """cn_block_idx = int(index_block / (n_base / n_ctrl)) (src/UniGenTransformer.py:1126-1127,1159-1160)."""
    interval = n_base / n_ctrl
    return [int(i / interval) for i in range(n_base)]


def weave(h, enc, temb, preprocess, base_double, base_single, ctrl_double, ctrl_single, add_double, add_single,
          conditioning_scale: float, method: str = "overall_add", rec=None):
    """UniGenFlux.base_forward + control_forward (src/UniGenTransformer.py:1070-1180) over pluggable blocks.
    base_double[i](h, enc, temb) -> (enc, h); base_single[i](x, temb) -> x; ctrl_* take (.., condition_temb);
    add_*[j](x) are the zero-linears; preprocess(h, enc) -> moe_output dict, run at the FIRST control call only."""
    rec = rec or (lambda *_: None)
    moe = None
    sched_d = weave_schedule(len(base_double), len(ctrl_double))
    for i, blk in enumerate(base_double):
        enc, h = blk(h, enc, temb)
        rec(f"double.{i}.base_hidden", h); rec(f"double.{i}.base_context", enc)
        j = sched_d[i]
        if moe is None:  # :1084-1089 — first call: control stream := expert_hidden + expert_cond
            moe = preprocess(h, enc)
            ctrl_in = moe["expert_hidden_states"] + moe["expert_condition_hidden_states"]
            rec("moe.ctrl_in", ctrl_in)
        else:            # later calls: the control block reads the BASE stream (1st positional arg, :1137)
            ctrl_in = h
        _, ch = ctrl_double[j](ctrl_in, moe["control_encoder_hidden_states"], moe["condition_temb"])
        h = h + add_double[j](ch) * conditioning_scale
        rec(f"double.{i}.ctrl_hidden", ch); rec(f"double.{i}.hidden", h)
    T = enc.shape[1]
    x = torch.cat([enc, h], dim=1)  # :1146 text first
    sched_s = weave_schedule(len(base_single), len(ctrl_single)) if len(ctrl_single) else []
    for i, blk in enumerate(base_single):
        x = blk(x, temb)
        rec(f"single.{i}.base_hidden", x)
        if len(ctrl_single):
            j = sched_s[i]
            cx = ctrl_single[j](x, moe["condition_temb"])
            zero = add_single[j](cx) * conditioning_scale
            if method == "overall_add":  # :1166-1172
                x = x + zero
            else:
                x = torch.cat([x[:, :T], x[:, T:] + zero[:, T:]], dim=1)
        rec(f"single.{i}.hidden", x)
    return x[:, T:], enc, moe


# --- src/lora_switching_module.py:4-39 + peft 0.15 LoraLayer.set_scale ---
def module_active_adapters(module) -> List[str]:
    if hasattr(module, "active_adapters"):
        return [a for a in module.active_adapters if a in module.scaling.keys()]
    return []


class enable_lora:
    """Context manager: zero the scale of every active adapter not in `enable_adapters`, restore on exit through
    `set_scale(adapter, saved_scaling)` (which re-multiplies by lora_alpha/r — exact only when alpha == r; replicated)."""

    def __init__(self, lora_modules, enable_adapters, tuner_type=None):
        self.lora_modules = [m for m in lora_modules if (tuner_type is None and hasattr(m, "set_scale"))
                             or (tuner_type is not None and isinstance(m, tuner_type))]
        self.saved = [{a: m.scaling[a] for a in module_active_adapters(m)} for m in self.lora_modules]
        self.enable_adapters = enable_adapters

    def __enter__(self):
        for m in self.lora_modules:
            for a in module_active_adapters(m):
                if a not in self.enable_adapters:
                    m.set_scale(a, 0)

    def __exit__(self, *exc):
        for i, m in enumerate(self.lora_modules):
            for a in module_active_adapters(m):
                m.set_scale(a, self.saved[i][a])


def lora_linear(x: Tensor, weight: Tensor, bias: Optional[Tensor], adapters: Dict[str, Tuple[Tensor, Tensor, float]],
                active: List[str]) -> Tensor:
    """peft 0.15 lora.Linear.forward: base(x) + sum_a B_a(A_a(x)) * scaling_a over active adapters."""
    y = F.linear(x, weight, bias)
    for a in active:
        A, Bm, scaling = adapters[a]
        y = y + F.linear(F.linear(x, A), Bm) * scaling
    return y


def segment_mask(bounds: List[int], visible: List[int]) -> Tensor:
    """Dense boolean (S,S) attention mask from the segment rule: query in segment i attends keys of
    segment j iff bit j of visible[i]. P-variant rule: txt,img -> all; c_i -> {txt, img, c_i}."""
    S = bounds[-1]
    seg = torch.zeros(S, dtype=torch.long)
    for i in range(len(bounds) - 1):
        seg[bounds[i]:bounds[i + 1]] = i
    n = len(visible)
    vm = torch.tensor([[(visible[i] >> j) & 1 for j in range(n)] for i in range(n)], dtype=torch.bool)
    return vm[seg][:, seg]


def pvariant_visibility(n_cond: int, strict: bool = False) -> List[int]:
    """Segments [txt, img, c_1..c_n]. Reference rule (UniCombineTransformerBlock.pyc L98-110): main queries see every
    segment; c_i sees txt, img, c_i. strict=True is the north-star wording (condition tokens see only themselves)."""
    n = 2 + n_cond
    vis = [(1 << n) - 1, (1 << n) - 1]
    for i in range(n_cond):
        vis.append((1 << (2 + i)) | (0 if strict else 0b11))
    return vis


# ------------------------------------------------------------------------------------------------------------------
# UniGenFlux forward (S-variant)
# ------------------------------------------------------------------------------------------------------------------
class UniGenFluxOracle:
    """Functional restatement of UniGenFlux (src/UniGenTransformer.py:712-1271) over a reference-keyed state dict."""

    def __init__(self, cfg: FluxConfig, state_dict: Dict[str, Tensor]):
        self.cfg = cfg
        self.sd = state_dict
        self.trace: Dict[str, Tensor] = {}
        self.record = False

    def _rec(self, name: str, t: Tensor) -> None:
        if self.record:
            self.trace[name] = t.detach().clone()

    # --- src/UniGenTransformer.py:925-967 (modulated branch, taken because use_rope) ---
    def expert_forward(self, hidden, cond, pooled, cond_pooled) -> Tuple[Tensor, Tensor]:
        """Inputs are the dispatched (1,E,C,·) tensors; returns stacked (1,E,C,D) outputs."""
        E = self.cfg.expert_nums
        outs_h, outs_c = [], []
        for e in range(E):
            p = f"moe.moe_layer.experts.deepspeed_experts.{e}"
            hc, cc = hidden[:, e], cond[:, e]
            s_c = linear(self.sd, f"{p}.0.1", cond_pooled[:, e])
            cc = modulated_flatten(cc, self.sd[f"{p}.0.0.weight"], s_c) + self.sd[f"{p}.0.0.bias"][None]
            s_h = linear(self.sd, f"{p}.1.1", pooled[:, e])
            hc = modulated_flatten(hc + cc, self.sd[f"{p}.1.0.weight"], s_h) + self.sd[f"{p}.1.0.bias"][None]
            outs_h.append(hc)
            outs_c.append(cc)
        return torch.stack(outs_h, dim=1), torch.stack(outs_c, dim=1)

    # --- src/UniGenUtils.py:74-134 + src/UniGenTransformer.py:969-1026 ---
    def moe_forward(self, hidden, cond, enc_ctrl, temb_ctrl, cond_temb, pooled, cond_pooled, ids, rts_uniform):
        cfg, sd = self.cfg, self.sd
        B, N, D = hidden.shape
        E = cfg.expert_nums
        choice = hidden + cond
        reshaped = choice.reshape(-1, D)
        logits = F.linear(reshaped.float(), sd["moe.moe_layer.gate.wg.weight"].float())
        C = moe_capacity(reshaped.shape[0], E)
        l_aux, combine, dispatch, exp_counts, sparse = top1gating(logits, C, rts_uniform)
        self._rec("moe.expert_idx", sparse[0]); self._rec("moe.slot", sparse[1]); self._rec("moe.prob", sparse[2])

        def disp(v):  # src/UniGenUtils.py:104-120
            if v.dim() == 2:
                v = v[:, None, :].expand(-1, N, -1).reshape(-1, v.shape[-1])
            else:
                v = v.reshape(-1, v.shape[-1])
            return moe_dispatch(dispatch, v)[None]  # (1,E,C,c)

        eh, ec = self.expert_forward(disp(hidden), disp(cond), disp(pooled), disp(cond_pooled))
        expert_hidden = moe_combine(combine, eh.reshape(E, C, D), choice)
        expert_cond = moe_combine(combine, ec.reshape(E, C, D), choice)
        self._rec("moe.expert_hidden", expert_hidden); self._rec("moe.expert_cond", expert_cond)

        expert_hidden, expert_cond = self.post_experts(hidden, cond, expert_hidden, expert_cond, enc_ctrl, temb_ctrl, cond_temb, ids)
        return expert_hidden, expert_cond, l_aux, exp_counts

    # --- src/UniGenTransformer.py:982-1024: what moe_forward does with the routed experts' (combined) outputs ---
    def post_experts(self, hidden, cond, expert_hidden, expert_cond, enc_ctrl, temb_ctrl, cond_temb, ids):
        cfg, sd = self.cfg, self.sd
        N = hidden.shape[1]
        H = cfg.num_attention_heads
        img_ids, txt_ids, cond_ids = ids
        rope_of = lambda *parts: flux_pos_embed(torch.cat(parts, 0), cfg.axes_dims_rope, cfg.theta)  # noqa: E731 (encoder ids first)
        # `expert_output` is the tuple the MoE layer returned (:979); the consis branch only rebinds the two LOCAL names (:1002-1003)
        # and only the shared-expert branch builds a new tuple from them (:1024). So with use_shared_expert=False the consis
        # module's result never reaches the return value (:1026) — restated as is (pinned by tests/golden: moe_wiring).
        expert_output = (expert_hidden, expert_cond)
        if cfg.use_consis_module and cfg.use_shared_expert:  # :982-1003, V2 — consis_module[0] is called TWICE, consis_module[1] never
            # hidden = the experts' condition output, encoder = the embedded condition tokens, temb = THIS condition's temb
            _, consis_cond = flux_double_block(sd, "consis_module.0", H, expert_cond, cond, cond_temb, rope_of(cond_ids, cond_ids))
            # hidden = [experts' image output | the block's condition output], encoder = the image tokens, temb = control temb
            _, hc = flux_double_block(sd, "consis_module.0", H, torch.cat([expert_hidden, consis_cond], dim=1), hidden, temb_ctrl,
                                      rope_of(img_ids, img_ids, cond_ids))
            self._rec("moe.consis_hidden", hc[:, :N]); self._rec("moe.consis_cond", hc[:, N:])
            expert_hidden, expert_cond = expert_hidden + hc[:, :N], expert_cond + hc[:, N:]
        if cfg.use_shared_expert:  # :1005-1024, V2
            cond_states, hid = flux_double_block(sd, "shared_expert.0", H, hidden, cond, cond_temb, rope_of(cond_ids, img_ids))
            hc = torch.cat([hid, cond_states], dim=1)
            _, hc = flux_double_block(sd, "shared_expert.1", H, hc, enc_ctrl, temb_ctrl, rope_of(txt_ids, img_ids, cond_ids))
            hid, cond_states = hc[:, :N], hc[:, N:]
            self._rec("moe.shared_hidden", hid); self._rec("moe.shared_cond", cond_states)
            expert_output = (hid + expert_hidden, cond_states + expert_cond)
        return expert_output

    # --- src/UniGenTransformer.py:1028-1068 ---
    def preprocess_moe_forward(self, hidden, cond_tokens, enc, pooled, cond_pooled, timestep, guidance, ids,
                               rts_uniform):
        cfg, sd = self.cfg, self.sd
        cond = linear(sd, "control_x_embedder", cond_tokens)
        ctrl_pooled = pooled if cfg.use_pooled_prompt_embeds else torch.zeros_like(pooled)
        control_temb = combined_timestep_text_embed(sd, "control_time_text_embed", timestep, ctrl_pooled, guidance)
        condition_temb = combined_timestep_text_embed(sd, "control_condition_embed", timestep, cond_pooled, guidance)
        enc_ctrl = linear(sd, "control_context_embedder", enc)
        eh, ec, l_aux, exp_counts = self.moe_forward(hidden, cond, enc_ctrl, control_temb, condition_temb, pooled,
                                                     cond_pooled, ids, rts_uniform)
        return dict(expert_hidden_states=eh, expert_condition_hidden_states=ec,
                    control_encoder_hidden_states=enc_ctrl, control_temb=control_temb,
                    condition_temb=condition_temb, exp_count=exp_counts, moe_loss=l_aux)

    # --- src/UniGenTransformer.py:1182-1271 (+ base_forward :1106-1180, control_forward :1070-1104) ---
    def forward(self, hidden_states, condition_hidden_states, conditioning_scale=1.0, encoder_hidden_states=None,
                pooled_projections=None, condition_pooled_projections=None, timestep=None, img_ids=None,
                txt_ids=None, guidance=None, condition_ids=None, rts_uniform=None):
        cfg, sd = self.cfg, self.sd
        H = cfg.num_attention_heads
        h = linear(sd, "x_embedder", hidden_states)
        if guidance is not None:
            guidance = guidance.to(h.dtype) * 1000
        timestep = timestep.to(h.dtype) * 1000
        temb = combined_timestep_text_embed(sd, "time_text_embed", timestep, pooled_projections, guidance)
        enc = linear(sd, "context_embedder", encoder_hidden_states)
        if txt_ids.dim() == 3:
            txt_ids = txt_ids[0]
        if img_ids.dim() == 3:
            img_ids = img_ids[0]
        ids = torch.cat((txt_ids, img_ids), dim=0)
        rope = flux_pos_embed(ids, cfg.axes_dims_rope, cfg.theta)
        rope_ctrl = rope  # control_pos_embed_input is a deepcopy of the parameter-free pos_embed (:727), same ids order
        self._rec("temb", temb); self._rec("x_embed", h); self._rec("context_embed", enc)

        def preprocess(h_, enc_):
            if isinstance(condition_hidden_states, (list, tuple)):
                # MultiCondtionUniGenFlux.preprocess_moe_forward (src/UniGenTransformer.py:1275-1322): one full CoMoE
                # pass per condition; control stream := sum_c (expert_hidden + expert_cond); condition_temb := sum_c;
                # moe_loss / exp_count are those of the LAST condition (:1320-1321).
                merged, merged_temb, last = 0, 0, None
                for c, (cid, chs, cp) in enumerate(zip(condition_ids, condition_hidden_states,
                                                       condition_pooled_projections)):
                    cid = cid[0] if cid.dim() == 3 else cid
                    last = self.preprocess_moe_forward(h_, chs, enc_, pooled_projections, cp, timestep, guidance,
                                                       (img_ids, txt_ids, cid), rts_uniform[c])
                    merged = merged + (last["expert_hidden_states"] + last["expert_condition_hidden_states"])
                    merged_temb = merged_temb + last["condition_temb"]
                    self._rec(f"moe.cond{c}.ctrl_in", last["expert_hidden_states"] + last["expert_condition_hidden_states"])
                z = torch.zeros_like(merged)
                return dict(expert_hidden_states=merged, expert_condition_hidden_states=z,
                            control_encoder_hidden_states=last["control_encoder_hidden_states"],
                            control_temb=last["control_temb"], condition_temb=merged_temb,
                            exp_count=last["exp_count"], moe_loss=last["moe_loss"])
            return self.preprocess_moe_forward(h_, condition_hidden_states, enc_, pooled_projections,
                                               condition_pooled_projections, timestep, guidance,
                                               (img_ids, txt_ids, condition_ids), rts_uniform)

        dbl = lambda p: (lambda h_, c_, t_: flux_double_block(sd, p, H, h_, c_, t_, rope))  # noqa: E731
        sgl = lambda p: (lambda x_, t_: flux_single_block(sd, p, H, x_, t_, rope_ctrl))      # noqa: E731
        lin = lambda p: (lambda x_: linear(sd, p, x_))                                        # noqa: E731
        h, enc, moe = weave(
            h, enc, temb, preprocess,
            [dbl(f"transformer_blocks.{i}") for i in range(cfg.num_layers)],
            [sgl(f"single_transformer_blocks.{i}") for i in range(cfg.num_single_layers)],
            [dbl(f"control_joint_trans_blocks.{j}") for j in range(cfg.cn_joint_layers)],
            [sgl(f"control_single_trans_blocks.{j}") for j in range(cfg.cn_single_layers)],
            [lin(f"controlnet_add_joint_blocks.{j}") for j in range(cfg.cn_joint_layers)],
            [lin(f"controlnet_add_single_blocks.{j}") for j in range(cfg.cn_single_layers)],
            conditioning_scale, cfg.single_block_control_method, self._rec)
        out = linear(sd, "proj_out", ada_layer_norm_continuous(sd, "norm_out", h, temb))
        self._rec("velocity", out)
        return out, dict(moe_loss=moe["moe_loss"] * 0.1), dict(expert_counts=moe["exp_count"])


# ------------------------------------------------------------------------------------------------------------------
# deterministic random-init weights with the reference's state-dict names
# ------------------------------------------------------------------------------------------------------------------
def _lin(sd, name, out_f, in_f, gen, zero_linear_std=None):
    """nn.Linear default init (kaiming_uniform(a=sqrt(5)) -> U(-1/sqrt(fan_in), 1/sqrt(fan_in)) for both)."""
    if zero_linear_std is not None:
        sd[name + ".weight"] = torch.randn(out_f, in_f, generator=gen) * zero_linear_std
        sd[name + ".bias"] = torch.randn(out_f, generator=gen) * zero_linear_std
        return
    bound = 1.0 / math.sqrt(in_f)
    sd[name + ".weight"] = (torch.rand(out_f, in_f, generator=gen) * 2 - 1) * bound
    sd[name + ".bias"] = (torch.rand(out_f, generator=gen) * 2 - 1) * bound


def _double_block(sd, p, D, dh, gen):
    _lin(sd, p + ".norm1.linear", 6 * D, D, gen)
    _lin(sd, p + ".norm1_context.linear", 6 * D, D, gen)
    for n in ("to_q", "to_k", "to_v", "add_q_proj", "add_k_proj", "add_v_proj", "to_out.0", "to_add_out"):
        _lin(sd, f"{p}.attn.{n}", D, D, gen)
    for n in ("norm_q", "norm_k", "norm_added_q", "norm_added_k"):
        sd[f"{p}.attn.{n}.weight"] = torch.ones(dh) + 0.1 * torch.randn(dh, generator=gen)
    for ff in ("ff", "ff_context"):
        _lin(sd, f"{p}.{ff}.net.0.proj", 4 * D, D, gen)
        _lin(sd, f"{p}.{ff}.net.2", D, 4 * D, gen)


def _single_block(sd, p, D, dh, gen):
    _lin(sd, p + ".norm.linear", 3 * D, D, gen)
    _lin(sd, p + ".proj_mlp", 4 * D, D, gen)
    _lin(sd, p + ".proj_out", D, 5 * D, gen)
    for n in ("to_q", "to_k", "to_v"):
        _lin(sd, f"{p}.attn.{n}", D, D, gen)
    for n in ("norm_q", "norm_k"):
        sd[f"{p}.attn.{n}.weight"] = torch.ones(dh) + 0.1 * torch.randn(dh, generator=gen)


def _time_text(sd, p, D, pooled_dim, gen, guidance):
    _lin(sd, p + ".timestep_embedder.linear_1", D, 256, gen)
    _lin(sd, p + ".timestep_embedder.linear_2", D, D, gen)
    if guidance:
        _lin(sd, p + ".guidance_embedder.linear_1", D, 256, gen)
        _lin(sd, p + ".guidance_embedder.linear_2", D, D, gen)
    _lin(sd, p + ".text_embedder.linear_1", D, pooled_dim, gen)
    _lin(sd, p + ".text_embedder.linear_2", D, D, gen)


def init_state_dict(cfg: FluxConfig, seed: int = 0, zero_linear_std: Optional[float] = 0.02) -> Dict[str, Tensor]:
    """Random-init weights under the reference's names. Zero-linears (`zero_module`, src/UniGenUtils.py:194-197) get
    N(0, zero_linear_std) so the control branch is visible to parity; pass None for true zeros. RMSNorm weights are
    1 + 0.1 N(0,1) (instead of the default all-ones) so a missing / misplaced norm weight is detectable."""
    gen = torch.Generator().manual_seed(seed)
    D, dh = cfg.inner_dim, cfg.attention_head_dim
    sd: Dict[str, Tensor] = {}
    _lin(sd, "x_embedder", D, cfg.in_channels, gen)
    _lin(sd, "context_embedder", D, cfg.joint_attention_dim, gen)
    _time_text(sd, "time_text_embed", D, cfg.pooled_projection_dim, gen, cfg.guidance_embeds)
    for i in range(cfg.num_layers):
        _double_block(sd, f"transformer_blocks.{i}", D, dh, gen)
    for i in range(cfg.num_single_layers):
        _single_block(sd, f"single_transformer_blocks.{i}", D, dh, gen)
    _lin(sd, "norm_out.linear", 2 * D, D, gen)
    _lin(sd, "proj_out", cfg.in_channels, D, gen)
    # control branch (src/UniGenTransformer.py:727-773)
    _time_text(sd, "control_time_text_embed", D, cfg.pooled_projection_dim, gen, cfg.guidance_embeds)
    _time_text(sd, "control_condition_embed", D, cfg.pooled_projection_dim, gen, cfg.guidance_embeds)
    _lin(sd, "control_context_embedder", D, D, gen)
    _lin(sd, "control_x_embedder", D, cfg.in_channels, gen)
    for j in range(cfg.cn_joint_layers):
        _double_block(sd, f"control_joint_trans_blocks.{j}", D, dh, gen)
        if zero_linear_std is None:
            sd[f"controlnet_add_joint_blocks.{j}.weight"] = torch.zeros(D, D)
            sd[f"controlnet_add_joint_blocks.{j}.bias"] = torch.zeros(D)
        else:
            _lin(sd, f"controlnet_add_joint_blocks.{j}", D, D, gen, zero_linear_std)
    for j in range(cfg.cn_single_layers):
        _single_block(sd, f"control_single_trans_blocks.{j}", D, dh, gen)
        if zero_linear_std is None:
            sd[f"controlnet_add_single_blocks.{j}.weight"] = torch.zeros(D, D)
            sd[f"controlnet_add_single_blocks.{j}.bias"] = torch.zeros(D)
        else:
            _lin(sd, f"controlnet_add_single_blocks.{j}", D, D, gen, zero_linear_std)
    # CoMoE (src/UniGenTransformer.py:806-891): each expert individually seeded (SURVEY.md §A.5 note)
    sd["moe.moe_layer.gate.wg.weight"] = (torch.rand(cfg.expert_nums, D, generator=gen) * 2 - 1) / math.sqrt(D)
    for e in range(cfg.expert_nums):
        p = f"moe.moe_layer.experts.deepspeed_experts.{e}"
        for br in (0, 1):
            _lin(sd, f"{p}.{br}.0", D, D, gen)
            _lin(sd, f"{p}.{br}.1", D, cfg.pooled_projection_dim, gen)
    for s in (0, 1):
        _double_block(sd, f"shared_expert.{s}", D, dh, gen)
    if cfg.use_consis_module:  # both exist in the reference's state dict; only consis_module.0 is ever called (:994, :998)
        for s in (0, 1):
            _double_block(sd, f"consis_module.{s}", D, dh, gen)
    return sd


def make_inputs(cfg: FluxConfig, height: int, width: int, text_len: int = 512, batch: int = 1, seed: int = 1234,
                step: int = 0, steps: int = 4, condition_type: str = "canny") -> Dict[str, Tensor]:
    """Synthetic inputs of SURVEY.md §8(d): N = (H/16)(W/16) image tokens, Nc = N condition tokens, T text tokens."""
    gen = torch.Generator().manual_seed(seed)
    h2, w2 = height // 16, width // 16
    N = h2 * w2
    img_ids = prepare_latent_image_ids(h2, w2)
    cond_ids, _ = condition_ids(condition_type, height, width)
    ts = torch.linspace(1, 1 / steps, steps)[step]
    return dict(
        hidden_states=torch.randn(batch, N, cfg.in_channels, generator=gen),
        condition_hidden_states=torch.randn(batch, N, cfg.in_channels, generator=gen),
        encoder_hidden_states=torch.randn(batch, text_len, cfg.joint_attention_dim, generator=gen),
        pooled_projections=torch.randn(batch, cfg.pooled_projection_dim, generator=gen),
        condition_pooled_projections=torch.randn(batch, cfg.pooled_projection_dim, generator=gen),
        timestep=ts.expand(batch).clone(),
        img_ids=img_ids, txt_ids=torch.zeros(text_len, 3), condition_ids=cond_ids,
        rts_uniform=torch.rand(batch * N, cfg.expert_nums, generator=gen),
    )


def make_multi_inputs(cfg: FluxConfig, height: int, width: int, condition_types=("depth", "canny", "subject"),
                      text_len: int = 512, batch: int = 1, seed: int = 1234, step: int = 0, steps: int = 4):
    """Inputs for MultiCondtionUniGenFlux (src/UniGenTransformer.py:1360-1450): LISTS of condition tokens / pooled
    embeddings / ids, one RTS uniform draw per condition (every MoE call draws its own)."""
    base = make_inputs(cfg, height, width, text_len, batch, seed, step, steps, condition_types[0])
    gen = torch.Generator().manual_seed(seed + 1)
    N = base["hidden_states"].shape[1]
    base["condition_hidden_states"] = [torch.randn(batch, N, cfg.in_channels, generator=gen) for _ in condition_types]
    base["condition_pooled_projections"] = [torch.randn(batch, cfg.pooled_projection_dim, generator=gen)
                                            for _ in condition_types]
    base["condition_ids"] = [condition_ids(t, height, width)[0] for t in condition_types]
    base["rts_uniform"] = [torch.rand(batch * N, cfg.expert_nums, generator=gen) for _ in condition_types]
    return base


# ------------------------------------------------------------------------------------------------------------------
# P-variant (predecessor bytecode): LoRA-switched joint blocks where the condition tokens run through every block
# (UniCombineTransformerBlock.pyc attn_forward L9-136, block_forward L140-236,
# single_block_forward L239-295; UniCombineTransformer2DModel.pyc forward L53-211).  State dict = diffusers Flux names
# + PEFT names `<linear>.lora_A.<adapter>.weight` [r, in], `<linear>.lora_B.<adapter>.weight` [out, r].
# ------------------------------------------------------------------------------------------------------------------
PV_LORA_DOUBLE = ("norm1.linear", "attn.to_q", "attn.to_k", "attn.to_v", "attn.to_out.0", "ff.net.2")
PV_LORA_SINGLE = ("norm.linear", "proj_mlp", "proj_out", "attn.to_q", "attn.to_k", "attn.to_v")


def pv_lora_targets(cfg: FluxConfig) -> List[str]:
    names = ["x_embedder"]
    names += [f"transformer_blocks.{i}.{n}" for i in range(cfg.num_layers) for n in PV_LORA_DOUBLE]
    names += [f"single_transformer_blocks.{i}.{n}" for i in range(cfg.num_single_layers) for n in PV_LORA_SINGLE]
    return names


def init_pvariant_state_dict(cfg: FluxConfig, adapters: List[str], rank: int = 4, seed: int = 0) -> Dict[str, Tensor]:
    """Base Flux weights + one LoRA pair per adapter on every switched linear (lora_B is NOT zero-initialised so the
    adapters are visible to parity)."""
    gen = torch.Generator().manual_seed(seed)
    D, dh = cfg.inner_dim, cfg.attention_head_dim
    sd: Dict[str, Tensor] = {}
    _lin(sd, "x_embedder", D, cfg.in_channels, gen)
    _lin(sd, "context_embedder", D, cfg.joint_attention_dim, gen)
    _time_text(sd, "time_text_embed", D, cfg.pooled_projection_dim, gen, cfg.guidance_embeds)
    for i in range(cfg.num_layers):
        _double_block(sd, f"transformer_blocks.{i}", D, dh, gen)
    for i in range(cfg.num_single_layers):
        _single_block(sd, f"single_transformer_blocks.{i}", D, dh, gen)
    _lin(sd, "norm_out.linear", 2 * D, D, gen)
    _lin(sd, "proj_out", cfg.in_channels, D, gen)
    for name in pv_lora_targets(cfg):
        out_f, in_f = sd[name + ".weight"].shape
        for a in adapters:
            sd[f"{name}.lora_A.{a}.weight"] = torch.randn(rank, in_f, generator=gen) / math.sqrt(in_f)
            sd[f"{name}.lora_B.{a}.weight"] = torch.randn(out_f, rank, generator=gen) * (0.5 / math.sqrt(rank))
    return sd


class PVariantOracle:
    """Functional restatement of the predecessor's UniCombineTransformer2DModel forward with `enable_lora` switching.
    adapters: every adapter loaded on the model; condition_types: the adapter named after each condition;
    den = adapters - condition_types act on the text/image rows. scaling[a] = lora_alpha/r (1.0 when alpha == r)."""

    def __init__(self, cfg: FluxConfig, sd: Dict[str, Tensor], adapters: List[str], scaling: Optional[Dict[str, float]] = None,
                 strict_mask: bool = False, add_cond_attn: bool = False):
        self.cfg, self.sd, self.adapters = cfg, sd, list(adapters)
        self.scaling = scaling or {a: 1.0 for a in adapters}
        self.strict_mask = strict_mask
        self.add_cond_attn = add_cond_attn  # model_config['add_cond_attn'] (UniCombineTransformerBlock.pyc L201-202)
        self.trace: Dict[str, Tensor] = {}
        self.record = False

    def _rec(self, k, v):
        if self.record:
            self.trace[k] = v.detach().clone()

    def lin(self, name: str, x: Tensor, active: List[str]) -> Tensor:
        """peft lora.Linear.forward under `enable_lora([module], active)`: adapters outside `active` have scale 0."""
        y = F.linear(x, self.sd[name + ".weight"], self.sd.get(name + ".bias"))
        for a in active:
            ka = f"{name}.lora_A.{a}.weight"
            if ka in self.sd:
                y = y + F.linear(F.linear(x, self.sd[ka]), self.sd[f"{name}.lora_B.{a}.weight"]) * self.scaling[a]
        return y

    def forward(self, hidden_states, condition_latents, condition_ids, condition_types, encoder_hidden_states,
                pooled_projections, timestep, img_ids, txt_ids, c_t: float = 0.0, return_condition_latents: bool = False):
        cfg, sd = self.cfg, self.sd
        H = cfg.num_attention_heads
        den = [a for a in self.adapters if a not in condition_types]
        sets = [[t] for t in condition_types]
        n = len(condition_types)
        # 2DModel L92-99: x_embedder under enable_lora per segment
        h = self.lin("x_embedder", hidden_states, den)
        conds = [self.lin("x_embedder", c, sets[i]) for i, c in enumerate(condition_latents)]
        timestep = timestep.to(h.dtype) * 1000
        temb = combined_timestep_text_embed(sd, "time_text_embed", timestep, pooled_projections)
        cond_temb = combined_timestep_text_embed(sd, "time_text_embed", torch.ones_like(timestep) * c_t * 1000,
                                                 pooled_projections)  # L118-121
        enc = linear(sd, "context_embedder", encoder_hidden_states)
        rope = flux_pos_embed(torch.cat([txt_ids, img_ids], 0), cfg.axes_dims_rope, cfg.theta)
        cond_ropes = [flux_pos_embed(ci, cfg.axes_dims_rope, cfg.theta) for ci in condition_ids]
        T, N = enc.shape[1], h.shape[1]
        bounds = [0, T, T + N]
        for c in conds:
            bounds.append(bounds[-1] + c.shape[1])
        mask = segment_mask(bounds, pvariant_visibility(n, strict=self.strict_mask))
        rope_all = (torch.cat([rope[0]] + [r[0] for r in cond_ropes], 0), torch.cat([rope[1]] + [r[1] for r in cond_ropes], 0))

        def attention(q_segs, k_segs, v_segs):
            """Each query segment attends the keys its visibility row allows == one masked SDPA over the joint sequence
            (pyc L98-110: main queries vs [txt|img|c_1..c_n]; c_i queries vs [txt|img|c_i])."""
            Q, K, V = torch.cat(q_segs, 2), torch.cat(k_segs, 2), torch.cat(v_segs, 2)
            Q, K = apply_rotary_emb(Q, rope_all), apply_rotary_emb(K, rope_all)
            O = sdpa(Q, K, V, mask)
            B, _, S, dh = O.shape
            return O.transpose(1, 2).reshape(B, S, H * dh)

        for i in range(cfg.num_layers):  # ---- block_forward (pyc L140-236) ----
            p = f"transformer_blocks.{i}"
            e = self.lin(p + ".norm1.linear", F.silu(temb), den).chunk(6, dim=1)
            ec = [self.lin(p + ".norm1.linear", F.silu(cond_temb), s).chunk(6, dim=1) for s in sets]
            et = linear(sd, p + ".norm1_context.linear", F.silu(temb)).chunk(6, dim=1)
            nh = layer_norm(h) * (1 + e[1][:, None]) + e[0][:, None]
            nt = layer_norm(enc) * (1 + et[1][:, None]) + et[0][:, None]
            ncs = [layer_norm(c) * (1 + ec[j][1][:, None]) + ec[j][0][:, None] for j, c in enumerate(conds)]
            a = p + ".attn"
            hd = lambda x: _heads(x, H)  # noqa: E731
            q = [rms_norm(hd(linear(sd, a + ".add_q_proj", nt)), sd[a + ".norm_added_q.weight"]),
                 rms_norm(hd(self.lin(a + ".to_q", nh, den)), sd[a + ".norm_q.weight"])]
            k = [rms_norm(hd(linear(sd, a + ".add_k_proj", nt)), sd[a + ".norm_added_k.weight"]),
                 rms_norm(hd(self.lin(a + ".to_k", nh, den)), sd[a + ".norm_k.weight"])]
            v = [hd(linear(sd, a + ".add_v_proj", nt)), hd(self.lin(a + ".to_v", nh, den))]
            for j, nc in enumerate(ncs):  # same norm_q / norm_k weights for every segment (L34-37, L89-92)
                q.append(rms_norm(hd(self.lin(a + ".to_q", nc, sets[j])), sd[a + ".norm_q.weight"]))
                k.append(rms_norm(hd(self.lin(a + ".to_k", nc, sets[j])), sd[a + ".norm_k.weight"]))
                v.append(hd(self.lin(a + ".to_v", nc, sets[j])))
            O = attention(q, k, v)
            ot, oh = O[:, :T], O[:, T:T + N]
            h = h + e[2][:, None] * self.lin(a + ".to_out.0", oh, den)
            enc = enc + et[2][:, None] * linear(sd, a + ".to_add_out", ot)
            for j in range(n):  # pyc L198-202
                oc = O[:, bounds[2 + j]:bounds[3 + j]]
                gated = ec[j][2][:, None] * self.lin(a + ".to_out.0", oc, sets[j])
                conds[j] = conds[j] + gated
                if self.add_cond_attn:
                    h = h + gated
            # feed-forward: ff.net.2 is the switched linear (L222-230)
            def ff(x, sh, sc, gate, active):
                y = gelu_tanh(linear(sd, p + ".ff.net.0.proj", layer_norm(x) * (1 + sc[:, None]) + sh[:, None]))
                return x + gate[:, None] * self.lin(p + ".ff.net.2", y, active)
            h = ff(h, e[3], e[4], e[5], den)
            enc = enc + et[5][:, None] * feed_forward(sd, p + ".ff_context", layer_norm(enc) * (1 + et[4][:, None]) + et[3][:, None])
            conds = [ff(c, ec[j][3], ec[j][4], ec[j][5], sets[j]) for j, c in enumerate(conds)]
            self._rec(f"double.{i}.hidden", h); self._rec(f"double.{i}.context", enc)
            for j, c in enumerate(conds):
                self._rec(f"double.{i}.cond{j}", c)

        x = torch.cat([enc, h], dim=1)  # 2DModel L176
        for i in range(cfg.num_single_layers):  # ---- single_block_forward (pyc L239-295) ----
            p = f"single_transformer_blocks.{i}"
            e = self.lin(p + ".norm.linear", F.silu(temb), den).chunk(3, dim=1)
            ec = [self.lin(p + ".norm.linear", F.silu(cond_temb), s).chunk(3, dim=1) for s in sets]
            nx = layer_norm(x) * (1 + e[1][:, None]) + e[0][:, None]
            ncs = [layer_norm(c) * (1 + ec[j][1][:, None]) + ec[j][0][:, None] for j, c in enumerate(conds)]
            a = p + ".attn"
            hd = lambda t_: _heads(t_, H)  # noqa: E731
            segs = [(nx, den)] + [(nc, sets[j]) for j, nc in enumerate(ncs)]
            q = [rms_norm(hd(self.lin(a + ".to_q", s_, act)), sd[a + ".norm_q.weight"]) for s_, act in segs]
            k = [rms_norm(hd(self.lin(a + ".to_k", s_, act)), sd[a + ".norm_k.weight"]) for s_, act in segs]
            v = [hd(self.lin(a + ".to_v", s_, act)) for s_, act in segs]
            O = attention(q, k, v)
            mlps = [gelu_tanh(self.lin(p + ".proj_mlp", s_, act)) for s_, act in segs]
            x = x + e[2][:, None] * self.lin(p + ".proj_out", torch.cat([O[:, :T + N], mlps[0]], 2), den)
            for j in range(n):
                oc = O[:, bounds[2 + j]:bounds[3 + j]]
                conds[j] = conds[j] + ec[j][2][:, None] * self.lin(p + ".proj_out", torch.cat([oc, mlps[1 + j]], 2), sets[j])
            self._rec(f"single.{i}.hidden", x)
            for j, c in enumerate(conds):
                self._rec(f"single.{i}.cond{j}", c)
        h = x[:, T:]
        out = linear(sd, "proj_out", ada_layer_norm_continuous(sd, "norm_out", h, temb))
        self._rec("velocity", out)
        if return_condition_latents:  # 2DModel L203-209: the condition streams after the last single block, un-projected
            return out, conds
        return out


# ------------------------------------------------------------------------------------------------------------------
# denoise-loop glue (callers of the path)
# ------------------------------------------------------------------------------------------------------------------
def calculate_shift(image_seq_len, base_seq_len=256, max_seq_len=4096, base_shift=0.5, max_shift=1.15):
    """diffusers pipeline_flux.calculate_shift as called at src/UniGenPipeline.py:991-997: the reference passes
    `scheduler.config.get("max_shift", 1.15)` (and base 256 / 4096 / 0.5)."""
    m = (max_shift - base_shift) / (max_seq_len - base_seq_len)
    return image_seq_len * m + (base_shift - m * base_seq_len)


def flow_match_sigmas(n: int, image_seq_len: int, use_dynamic_shifting: bool = True) -> Tensor:
    """np.linspace(1, 1/n, n) (src/UniGenPipeline.py:989) + FlowMatchEulerDiscreteScheduler exponential time shift + [0]."""
    s = torch.linspace(1.0, 1.0 / n, n, dtype=torch.float64)
    if use_dynamic_shifting:
        mu = calculate_shift(image_seq_len)
        s = math.exp(mu) / (math.exp(mu) + (1.0 / s - 1.0))
    return torch.cat([s, torch.zeros(1, dtype=torch.float64)]).float()


def pack_latents(x: Tensor) -> Tensor:
    """FluxPipeline._pack_latents."""
    B, C, H, W = x.shape
    return x.view(B, C, H // 2, 2, W // 2, 2).permute(0, 2, 4, 1, 3, 5).reshape(B, (H // 2) * (W // 2), C * 4)


def unpack_latents(x: Tensor, H: int, W: int) -> Tensor:
    """FluxPipeline._unpack_latents with H, W the latent height / width."""
    B, _, C4 = x.shape
    return x.view(B, H // 2, W // 2, C4 // 4, 2, 2).permute(0, 3, 1, 4, 2, 5).reshape(B, C4 // 4, H, W)


def denoise_loop(model: "UniGenFluxOracle", inp: Dict[str, Tensor], steps: int, rts: List[Tensor]) -> Tensor:
    """Loop body of UniGenFLUXPipeline.__call__ (src/UniGenPipeline.py:1050-1116) with the Euler flow-match step."""
    x = inp["hidden_states"].clone()
    sig = flow_match_sigmas(steps, x.shape[1])
    for i in range(steps):
        args = dict(inp, hidden_states=x, timestep=sig[i].expand(x.shape[0]).clone(), rts_uniform=rts[i])
        v = model.forward(**args)[0]
        x = x + (sig[i + 1] - sig[i]) * v
    return x
