"""ORACLE — test infrastructure, NOT product code.

CPU-runnable, pure-torch restatement of the reference's SD3.5 denoiser forward: one `UniGenSD3.forward`
(src/UniGenTransformer.py:490-710 on top of `UniGenBase` :20-296) with the block overrides of
src/UniGenUtils.py:340-522 — SURVEY.md §8 row A16 / BASELINE cfg5. Only tests/, `__graft_entry__.smoke()` and
`bench.py`'s cpu_baseline / `--impl reference` legs may import this file.

Configuration restated: the SHIPPED control_params (config/unigen.yaml:3-11: `use_modulate: False`, no `use_rope`,
`use_shared_expert: True`, `use_encoder_hidden_states: True`, 3 experts per condition) — i.e. the experts are two
`SD3SingleTransformerBlock`s (src/UniGenTransformer.py:184-195) run on the dispatched `(1, C, D)` capacity buffers
with a PER-TOKEN `(1, C, D)` temb (src/UniGenUtils.py:375-414), the shared experts are
`JointTransformerBlock(context_pre_only=False)` and `JointTransformerBlock(context_pre_only=True, dual)` (:205-222).
Also restated: `use_modulate: True` (:171-183, :252-258 — the experts are the condition-modulated linear pairs of the
Flux class: `modulated_flatten` over `Linear(D, D)` weights scaled by `Linear(pooled_dim, D)` of the pooled
projections) and `use_shared_expert: False` (:203, :279 — the routed experts' outputs alone). NOT restated:
`use_rope: True` — unreachable as shipped: `UniGenSD3.forward` calls `prepare_latent_image_ids` (:657-659), a name
that is neither defined nor imported anywhere in the reference (`src/UniGenUtils.py` has no such function, the only
import of it, :2055, fails), so the first forward raises NameError.

Parity pinning status
  * reference-owned arithmetic — the three AdaLN forwards (src/UniGenUtils.py:340-373), `JointTransformerBlock.forward`
    (:440-522) and `SD3SingleTransformerBlock.forward` (:386-414) incl. their per-token-temb branches,
    `UniGenBase.expert_forward` / `moe_forward` (src/UniGenTransformer.py:222-296), the `UniGenSD3` weave
    (`base_forward` / `control_forward` / `preprocess_moe_forward` :498-623) and the un-patchify of `forward`
    (:693-704) — is PINNED: tests/golden/make_golden_sd3.py runs the real functions from /root/reference (stub modules
    for the absent third-party imports, stand-in sub-modules) and tests/test_oracle_sd3_golden.py replays the vectors.
  * third-party arithmetic (diffusers 0.32.2 `Attention` + `JointAttnProcessor2_0`, `PatchEmbed` and its sincos
    table, `FeedForward`, `RMSNorm`, `CombinedTimestepTextProjEmbeddings`, the stock `SD3Transformer2DModel` wiring;
    deepspeed 0.16.5 top1gating) is "parity unpinned" (packages absent here) and follows SURVEY.md §A.10.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

from .unigen_oracle import (_heads, _lin, _time_text, combined_timestep_text_embed, feed_forward, layer_norm, linear,
                            moe_capacity, moe_combine, moe_dispatch, modulated_flatten, rms_norm, sdpa, top1gating)

Tensor = torch.Tensor


@dataclass
class SD3Config:
    """diffusers SD3Transformer2DModel config (SD3.5-medium values) + UniGen control_params."""
    sample_size: int = 128
    patch_size: int = 2
    in_channels: int = 16
    out_channels: int = 16
    num_layers: int = 24
    attention_head_dim: int = 64
    num_attention_heads: int = 24
    joint_attention_dim: int = 4096
    pooled_projection_dim: int = 2048
    pos_embed_max_size: int = 384
    qk_norm: Optional[str] = "rms_norm"
    dual_attention_layers: Tuple[int, ...] = tuple(range(13))
    condition_nums: int = 1
    expert_num_each_condition: int = 3
    use_pooled_prompt_embeds: bool = True
    use_shared_expert: bool = True
    use_modulate: bool = False  # src/UniGenTransformer.py:171: modulated-linear experts instead of transformer blocks

    @property
    def inner_dim(self) -> int:
        return self.num_attention_heads * self.attention_head_dim

    @property
    def expert_nums(self) -> int:  # src/UniGenTransformer.py:164
        return (self.condition_nums + 1) * self.expert_num_each_condition

    @staticmethod
    def tiny() -> "SD3Config":
        """4 MMDiT blocks (2 with dual attention, last one context_pre_only), hidden 384 = 6 heads of 64."""
        return SD3Config(sample_size=32, num_layers=4, num_attention_heads=6, pos_embed_max_size=48,
                         dual_attention_layers=(0, 1))

    @staticmethod
    def medium() -> "SD3Config":
        return SD3Config()


# ------------------------------------------------------------------------------------------------------------------
# third-party pieces restated (diffusers 0.32.2) — parity unpinned
# ------------------------------------------------------------------------------------------------------------------
def _sincos_1d(embed_dim: int, pos: np.ndarray) -> np.ndarray:
    """get_1d_sincos_pos_embed_from_grid: [sin | cos] of pos x omega, float64."""
    omega = np.arange(embed_dim // 2, dtype=np.float64)
    omega /= embed_dim / 2.0
    omega = 1.0 / 10000 ** omega
    out = np.einsum("m,d->md", pos.reshape(-1), omega)
    return np.concatenate([np.sin(out), np.cos(out)], axis=1)


def sincos_pos_embed_2d(embed_dim: int, grid_size: int, base_size: int, interpolation_scale: float = 1.0) -> Tensor:
    """get_2d_sincos_pos_embed(embed_dim, grid_size, base_size=..., interpolation_scale=1): meshgrid with W first
    (the MAE convention diffusers keeps), -> fp32 [1, grid*grid, embed_dim] as registered by PatchEmbed."""
    gh = np.arange(grid_size, dtype=np.float32) / (grid_size / base_size) / interpolation_scale
    gw = np.arange(grid_size, dtype=np.float32) / (grid_size / base_size) / interpolation_scale
    grid = np.stack(np.meshgrid(gw, gh), axis=0).reshape([2, 1, grid_size, grid_size])
    emb = np.concatenate([_sincos_1d(embed_dim // 2, grid[0]), _sincos_1d(embed_dim // 2, grid[1])], axis=1)
    return torch.from_numpy(emb).float().unsqueeze(0)


def cropped_pos_embed(table: Tensor, max_size: int, h: int, w: int) -> Tensor:
    """PatchEmbed.cropped_pos_embed: centre crop of the [1, max*max, D] table -> [1, h*w, D]."""
    top, left = (max_size - h) // 2, (max_size - w) // 2
    t = table.reshape(1, max_size, max_size, -1)[:, top:top + h, left:left + w, :]
    return t.reshape(1, h * w, -1)


def patch_embed(sd, prefix: str, latent: Tensor, cfg: SD3Config) -> Tensor:
    """PatchEmbed.forward (pos_embed_max_size set, layer_norm=False, flatten=True): Conv2d(k=p, s=p) -> (B, N, D)
    + cropped sincos table. Called at src/UniGenTransformer.py:517,663."""
    p = cfg.patch_size
    h, w = latent.shape[-2] // p, latent.shape[-1] // p
    x = F.conv2d(latent, sd[prefix + ".proj.weight"], sd[prefix + ".proj.bias"], stride=p)
    x = x.flatten(2).transpose(1, 2)
    pos = cropped_pos_embed(sd[prefix + ".pos_embed"], cfg.pos_embed_max_size, h, w)
    return (x + pos.to(x.dtype)).to(x.dtype)


def joint_attention(sd, prefix: str, H: int, x: Tensor, ctx: Optional[Tensor], context_pre_only: Optional[bool]):
    """diffusers Attention + JointAttnProcessor2_0: SAMPLE-FIRST concat, RMSNorm on q/k (and added q/k) when the
    checkpoint has the weights, no RoPE, no mask. Returns (sample_out, context_out | None)."""
    q, k, v = (_heads(linear(sd, f"{prefix}.to_{n}", x), H) for n in "qkv")
    if prefix + ".norm_q.weight" in sd:
        q, k = rms_norm(q, sd[prefix + ".norm_q.weight"]), rms_norm(k, sd[prefix + ".norm_k.weight"])
    if ctx is not None:
        cq, ck, cv = (_heads(linear(sd, f"{prefix}.add_{n}_proj", ctx), H) for n in "qkv")
        if prefix + ".norm_added_q.weight" in sd:
            cq, ck = rms_norm(cq, sd[prefix + ".norm_added_q.weight"]), rms_norm(ck, sd[prefix + ".norm_added_k.weight"])
        q, k, v = torch.cat([q, cq], 2), torch.cat([k, ck], 2), torch.cat([v, cv], 2)
    O = sdpa(q, k, v)
    B, _, S, dh = O.shape
    O = O.transpose(1, 2).reshape(B, S, H * dh).to(x.dtype)
    if ctx is None:
        return linear(sd, prefix + ".to_out.0", O), None
    n = x.shape[1]
    so, co = O[:, :n], O[:, n:]
    co = linear(sd, prefix + ".to_add_out", co) if not context_pre_only else None
    return linear(sd, prefix + ".to_out.0", so), co


# ------------------------------------------------------------------------------------------------------------------
# reference-owned pieces restated (src/UniGenUtils.py)
# ------------------------------------------------------------------------------------------------------------------
def _bc(v: Tensor, x: Tensor) -> Tensor:
    """`v[:, None]` when the AdaLN vector is per-sample (2-D) and x is 3-D, `v` when it is per-token (3-D)."""
    return v if v.dim() == x.dim() else v[:, None]


def ada_norm_zero(sd, prefix: str, x: Tensor, emb: Tensor):
    """adanorm_forward src/UniGenUtils.py:354-363 (AdaLayerNormZero with per-token emb support)."""
    e = linear(sd, prefix + ".linear", F.silu(emb))
    shift_msa, scale_msa, gate_msa, shift_mlp, scale_mlp, gate_mlp = e.chunk(6, dim=-1)
    return layer_norm(x) * (1 + _bc(scale_msa, x)) + _bc(shift_msa, x), gate_msa, shift_mlp, scale_mlp, gate_mlp


def ada_norm_zero_x(sd, prefix: str, x: Tensor, emb: Tensor):
    """sd35adanormX_forward src/UniGenUtils.py:340-352 (SD35AdaLayerNormZeroX, 9 chunks)."""
    e = linear(sd, prefix + ".linear", F.silu(emb))
    (shift_msa, scale_msa, gate_msa, shift_mlp, scale_mlp, gate_mlp, shift_msa2, scale_msa2, gate_msa2) = e.chunk(9, dim=-1)
    n = layer_norm(x)
    return (n * (1 + _bc(scale_msa, x)) + _bc(shift_msa, x), gate_msa, shift_mlp, scale_mlp, gate_mlp,
            n * (1 + _bc(scale_msa2, x)) + _bc(shift_msa2, x), gate_msa2)


def ada_norm_continuous(sd, prefix: str, x: Tensor, emb: Tensor) -> Tensor:
    """adanormContinuous_forward src/UniGenUtils.py:365-373: scale FIRST, then shift; chunk over dim=1."""
    e = linear(sd, prefix + ".linear", F.silu(emb).to(x.dtype))
    scale, shift = torch.chunk(e, 2, dim=1)
    if x.dim() == e.dim():
        return layer_norm(x) * (1 + scale) + shift
    return layer_norm(x) * (1 + scale)[:, None, :] + shift[:, None, :]


def joint_block(sd, prefix: str, H: int, h: Tensor, c: Tensor, temb: Tensor, dual: bool, context_pre_only: bool):
    """JointTransformerBlock.forward src/UniGenUtils.py:440-522. Returns (encoder_hidden_states | None, hidden_states)."""
    if dual:
        nh, gate_msa, shift_mlp, scale_mlp, gate_mlp, nh2, gate_msa2 = ada_norm_zero_x(sd, prefix + ".norm1", h, temb)
    else:
        nh, gate_msa, shift_mlp, scale_mlp, gate_mlp = ada_norm_zero(sd, prefix + ".norm1", h, temb)
    if context_pre_only:
        nc = ada_norm_continuous(sd, prefix + ".norm1_context", c, temb)
    else:
        nc, c_gate_msa, c_shift_mlp, c_scale_mlp, c_gate_mlp = ada_norm_zero(sd, prefix + ".norm1_context", c, temb)
    at, ca = joint_attention(sd, prefix + ".attn", H, nh, nc, context_pre_only)
    h = h + _bc(gate_msa, h) * at
    if dual:
        at2, _ = joint_attention(sd, prefix + ".attn2", H, nh2, None, None)
        h = h + _bc(gate_msa2, h) * at2
    nh = layer_norm(h) * (1 + _bc(scale_mlp, h)) + _bc(shift_mlp, h)
    h = h + _bc(gate_mlp, h) * feed_forward(sd, prefix + ".ff", nh)
    if context_pre_only:
        return None, h
    c = c + _bc(c_gate_msa, c) * ca
    nc = layer_norm(c) * (1 + _bc(c_scale_mlp, c)) + _bc(c_shift_mlp, c)
    c = c + _bc(c_gate_mlp, c) * feed_forward(sd, prefix + ".ff_context", nc)
    return c, h


def sd3_single_block(sd, prefix: str, H: int, x: Tensor, temb: Tensor) -> Tensor:
    """SD3SingleTransformerBlock.forward src/UniGenUtils.py:386-414 (self-attention, no qk-norm: the stock diffusers
    block builds `Attention` without `qk_norm`)."""
    nx, gate_msa, shift_mlp, scale_mlp, gate_mlp = ada_norm_zero(sd, prefix + ".norm1", x, temb)
    at, _ = joint_attention(sd, prefix + ".attn", H, nx, None, None)
    x = x + _bc(gate_msa, x) * at
    nx = layer_norm(x) * (1 + _bc(scale_mlp, x)) + _bc(shift_mlp, x)
    return x + _bc(gate_mlp, x) * feed_forward(sd, prefix + ".ff", nx)


def unpatchify(x: Tensor, h: int, w: int, p: int, c: int) -> Tensor:
    """src/UniGenTransformer.py:693-704: (B, h*w, p*p*c) -> (B, c, h*p, w*p), token channels ordered (p, q, c)."""
    x = x.reshape(x.shape[0], h, w, p, p, c)
    x = torch.einsum("nhwpqc->nchpwq", x)
    return x.reshape(x.shape[0], c, h * p, w * p)


class UniGenSD3Oracle:
    """Functional restatement of UniGenSD3 over a reference-keyed state dict."""

    def __init__(self, cfg: SD3Config, state_dict: Dict[str, Tensor]):
        self.cfg, self.sd = cfg, state_dict
        self.trace: Dict[str, Tensor] = {}
        self.record = False

    def _rec(self, name: str, t: Tensor) -> None:
        if self.record:
            self.trace[name] = t.detach().clone()

    # --- src/UniGenTransformer.py:222-262, transformer-block branch (:256-258) ---
    def expert_forward(self, hidden, cond, temb, cond_temb, pooled=None, cond_pooled=None) -> Tuple[Tensor, Tensor]:
        """Dispatched (1,E,C,D) tensors in, stacked (1,E,C,D) out. Every expert sees its C capacity slots as ONE
        sequence of a batch-1 sample: empty slots (all-zero rows, zero temb) take part in the self-attention.
        `use_modulate` (:252-255): expert[0] modulates the condition rows by the dispatched condition pooled projection,
        expert[1] modulates (hidden + that result) by the dispatched pooled projection; temb is not used."""
        H = self.cfg.num_attention_heads
        outs_h, outs_c = [], []
        for e in range(self.cfg.expert_nums):
            p = f"moe.moe_layer.experts.deepspeed_experts.{e}"
            if self.cfg.use_modulate:
                s_c = linear(self.sd, f"{p}.0.1", cond_pooled[:, e])
                cc = modulated_flatten(cond[:, e], self.sd[f"{p}.0.0.weight"], s_c) + self.sd[f"{p}.0.0.bias"][None]
                s_h = linear(self.sd, f"{p}.1.1", pooled[:, e])
                hc = modulated_flatten(hidden[:, e] + cc, self.sd[f"{p}.1.0.weight"], s_h) + self.sd[f"{p}.1.0.bias"][None]
                outs_h.append(hc)
                outs_c.append(cc)
                continue
            outs_h.append(sd3_single_block(self.sd, p + ".0", H, hidden[:, e], temb[:, e]))
            outs_c.append(sd3_single_block(self.sd, p + ".1", H, cond[:, e], cond_temb[:, e]))
        return torch.stack(outs_h, dim=1), torch.stack(outs_c, dim=1)

    # --- src/UniGenUtils.py:74-134 + src/UniGenTransformer.py:264-296 ---
    def moe_forward(self, hidden, cond, enc_ctrl, temb_ctrl, cond_temb, rts_uniform, pooled=None, cond_pooled=None):
        cfg, sd = self.cfg, self.sd
        B, N, D = hidden.shape
        E, H = cfg.expert_nums, cfg.num_attention_heads
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
            return moe_dispatch(dispatch, v)[None]

        if cfg.use_modulate:  # the pooled projections are dispatched like every other 2-D kwarg (src/UniGenUtils.py:104-110)
            eh, ec = self.expert_forward(disp(hidden), disp(cond), None, None, disp(pooled), disp(cond_pooled))
        else:
            eh, ec = self.expert_forward(disp(hidden), disp(cond), disp(temb_ctrl), disp(cond_temb))
        expert_hidden = moe_combine(combine, eh.reshape(E, C, D), choice)
        expert_cond = moe_combine(combine, ec.reshape(E, C, D), choice)
        self._rec("moe.expert_hidden", expert_hidden); self._rec("moe.expert_cond", expert_cond)
        if cfg.use_shared_expert:  # :281-294
            cond_states, hid = joint_block(sd, "shared_expert.0", H, hidden, cond, cond_temb, False, False)
            hc = torch.cat([hid, cond_states], dim=1)
            _, hc = joint_block(sd, "shared_expert.1", H, hc, enc_ctrl, temb_ctrl, True, True)
            hid, cond_states = hc[:, :N], hc[:, N:]
            self._rec("moe.shared_hidden", hid); self._rec("moe.shared_cond", cond_states)
            expert_hidden, expert_cond = hid + expert_hidden, cond_states + expert_cond
        return expert_hidden, expert_cond, l_aux, exp_counts

    # --- src/UniGenTransformer.py:498-540 ---
    def preprocess_moe_forward(self, hidden, cond_latents, enc, pooled, cond_pooled, timestep, rts_uniform):
        cfg, sd = self.cfg, self.sd
        cond = patch_embed(sd, "control_pos_embed_input", cond_latents, cfg)
        ctrl_pooled = pooled if cfg.use_pooled_prompt_embeds else torch.zeros_like(pooled)
        control_temb = combined_timestep_text_embed(sd, "control_time_text_embed", timestep, ctrl_pooled)
        condition_temb = combined_timestep_text_embed(sd, "control_condition_embed", timestep, cond_pooled)
        enc_ctrl = linear(sd, "control_context_embedder", enc)  # Linear(D, D) on the BASE text stream (:493)
        self._rec("moe.cond_embed", cond); self._rec("moe.enc_ctrl", enc_ctrl)
        eh, ec, l_aux, exp_counts = self.moe_forward(hidden, cond, enc_ctrl, control_temb, condition_temb, rts_uniform,
                                                     pooled, cond_pooled)
        return dict(expert_hidden_states=eh, expert_condition_hidden_states=ec, control_encoder_hidden_states=enc_ctrl,
                    control_temb=control_temb, condition_temb=condition_temb, exp_count=exp_counts, moe_loss=l_aux)

    # --- src/UniGenTransformer.py:625-710 (+ base_forward :583-623, control_forward :542-581) ---
    def forward(self, hidden_states, condition_hidden_states, conditioning_scale=1.0, encoder_hidden_states=None,
                pooled_projections=None, condition_pooled_projections=None, timestep=None, rts_uniform=None):
        cfg, sd = self.cfg, self.sd
        H = cfg.num_attention_heads
        height, width = hidden_states.shape[-2:]
        h = patch_embed(sd, "pos_embed", hidden_states, cfg)
        temb = combined_timestep_text_embed(sd, "time_text_embed", timestep, pooled_projections)
        enc = linear(sd, "context_embedder", encoder_hidden_states)
        self._rec("temb", temb); self._rec("x_embed", h); self._rec("context_embed", enc)
        moe = None
        n_ctrl = cfg.num_layers  # control_blocks_num defaults to num_layers (:30)
        for i in range(cfg.num_layers):
            last = i == cfg.num_layers - 1
            enc, h = joint_block(sd, f"transformer_blocks.{i}", H, h, enc, temb, i in cfg.dual_attention_layers, last)
            self._rec(f"block.{i}.base_hidden", h)
            j = int(i / (cfg.num_layers / n_ctrl))
            if i == 0:  # :558-562
                moe = self.preprocess_moe_forward(h, condition_hidden_states, enc, pooled_projections,
                                                  condition_pooled_projections, timestep, rts_uniform)
                ctrl_in = moe["expert_hidden_states"] + moe["expert_condition_hidden_states"]
                self._rec("moe.ctrl_in", ctrl_in)
            else:
                ctrl_in = h
            _, ch = joint_block(sd, f"control_transformer_blocks.{j}", H, ctrl_in, moe["control_encoder_hidden_states"],
                                moe["condition_temb"], j in cfg.dual_attention_layers, False)
            h = h + linear(sd, f"controlnet_add_blocks.{j}", ch) * conditioning_scale
            self._rec(f"block.{i}.ctrl_hidden", ch); self._rec(f"block.{i}.hidden", h)
        out = linear(sd, "proj_out", ada_norm_continuous(sd, "norm_out", h, temb))
        self._rec("proj_out", out)
        p = cfg.patch_size
        out = unpatchify(out, height // p, width // p, p, cfg.out_channels)
        self._rec("velocity", out)
        return out, dict(moe_loss=moe["moe_loss"] * 0.1), dict(expert_counts=moe["exp_count"])


# ------------------------------------------------------------------------------------------------------------------
# deterministic random-init weights with the reference's state-dict names
# ------------------------------------------------------------------------------------------------------------------
def _rms(sd, name, dh, gen):
    sd[name + ".weight"] = torch.ones(dh) + 0.1 * torch.randn(dh, generator=gen)


def _attn(sd, p, D, dh, gen, added: bool, add_out: bool, qk_norm: bool):
    for n in ("to_q", "to_k", "to_v", "to_out.0"):
        _lin(sd, f"{p}.{n}", D, D, gen)
    if qk_norm:
        _rms(sd, p + ".norm_q", dh, gen); _rms(sd, p + ".norm_k", dh, gen)
    if added:
        for n in ("add_q_proj", "add_k_proj", "add_v_proj"):
            _lin(sd, f"{p}.{n}", D, D, gen)
        if add_out:
            _lin(sd, p + ".to_add_out", D, D, gen)
        if qk_norm:
            _rms(sd, p + ".norm_added_q", dh, gen); _rms(sd, p + ".norm_added_k", dh, gen)


def _ff(sd, p, D, gen):
    _lin(sd, p + ".net.0.proj", 4 * D, D, gen)
    _lin(sd, p + ".net.2", D, 4 * D, gen)


def _joint_block_init(sd, p, D, dh, gen, dual: bool, context_pre_only: bool, qk_norm: bool):
    _lin(sd, p + ".norm1.linear", (9 if dual else 6) * D, D, gen)
    _lin(sd, p + ".norm1_context.linear", (2 if context_pre_only else 6) * D, D, gen)
    _attn(sd, p + ".attn", D, dh, gen, True, not context_pre_only, qk_norm)
    if dual:
        _attn(sd, p + ".attn2", D, dh, gen, False, False, qk_norm)
    _ff(sd, p + ".ff", D, gen)
    if not context_pre_only:
        _ff(sd, p + ".ff_context", D, gen)


def _patch_embed_init(sd, p, cfg: SD3Config, gen):
    D, k = cfg.inner_dim, cfg.patch_size
    fan_in = cfg.in_channels * k * k
    bound = 1.0 / math.sqrt(fan_in)
    sd[p + ".proj.weight"] = (torch.rand(D, cfg.in_channels, k, k, generator=gen) * 2 - 1) * bound
    sd[p + ".proj.bias"] = (torch.rand(D, generator=gen) * 2 - 1) * bound
    sd[p + ".pos_embed"] = sincos_pos_embed_2d(D, cfg.pos_embed_max_size, cfg.sample_size // cfg.patch_size)


def init_state_dict(cfg: SD3Config, seed: int = 0, zero_linear_std: Optional[float] = 0.02) -> Dict[str, Tensor]:
    gen = torch.Generator().manual_seed(seed)
    D, dh = cfg.inner_dim, cfg.attention_head_dim
    qk = cfg.qk_norm is not None
    sd: Dict[str, Tensor] = {}
    _patch_embed_init(sd, "pos_embed", cfg, gen)
    _time_text(sd, "time_text_embed", D, cfg.pooled_projection_dim, gen, False)
    _lin(sd, "context_embedder", D, cfg.joint_attention_dim, gen)
    for i in range(cfg.num_layers):
        _joint_block_init(sd, f"transformer_blocks.{i}", D, dh, gen, i in cfg.dual_attention_layers,
                          i == cfg.num_layers - 1, qk)
    _lin(sd, "norm_out.linear", 2 * D, D, gen)
    _lin(sd, "proj_out", cfg.patch_size * cfg.patch_size * cfg.out_channels, D, gen)
    # control branch (src/UniGenTransformer.py:25-136, 490-496)
    _patch_embed_init(sd, "control_pos_embed_input", cfg, gen)
    _time_text(sd, "control_time_text_embed", D, cfg.pooled_projection_dim, gen, False)
    _time_text(sd, "control_condition_embed", D, cfg.pooled_projection_dim, gen, False)
    _lin(sd, "control_context_embedder", D, D, gen)
    for j in range(cfg.num_layers):
        _joint_block_init(sd, f"control_transformer_blocks.{j}", D, dh, gen, j in cfg.dual_attention_layers, False, qk)
        if zero_linear_std is None:
            sd[f"controlnet_add_blocks.{j}.weight"], sd[f"controlnet_add_blocks.{j}.bias"] = torch.zeros(D, D), torch.zeros(D)
        else:
            _lin(sd, f"controlnet_add_blocks.{j}", D, D, gen, zero_linear_std)
    sd["moe.moe_layer.gate.wg.weight"] = (torch.rand(cfg.expert_nums, D, generator=gen) * 2 - 1) / math.sqrt(D)
    for e in range(cfg.expert_nums):
        for br in (0, 1):
            p = f"moe.moe_layer.experts.deepspeed_experts.{e}.{br}"
            if cfg.use_modulate:  # ModuleList([Linear(D, D), Linear(pooled_dim, D)]) (:173-182)
                _lin(sd, p + ".0", D, D, gen)
                _lin(sd, p + ".1", D, cfg.pooled_projection_dim, gen)
                continue
            _lin(sd, p + ".norm1.linear", 6 * D, D, gen)
            _attn(sd, p + ".attn", D, dh, gen, False, False, False)
            _ff(sd, p + ".ff", D, gen)
    if cfg.use_shared_expert:
        _joint_block_init(sd, "shared_expert.0", D, dh, gen, False, False, qk)
        _joint_block_init(sd, "shared_expert.1", D, dh, gen, True, True, qk)
    return sd


def make_inputs(cfg: SD3Config, height: int, width: int, text_len: int = 333, batch: int = 1, seed: int = 1234,
                timestep: float = 500.0) -> Dict[str, Tensor]:
    """Synthetic inputs: latents (B, 16, H/8, W/8) for image and condition, text (B, T, 4096), pooled (B, 2048);
    the SD3 pipeline passes the RAW scheduler timestep (0-1000), not t/1000 (src/UniGenPipeline.py:382,396)."""
    gen = torch.Generator().manual_seed(seed)
    lh, lw = height // 8, width // 8
    N = (lh // cfg.patch_size) * (lw // cfg.patch_size)
    return dict(
        hidden_states=torch.randn(batch, cfg.in_channels, lh, lw, generator=gen),
        condition_hidden_states=torch.randn(batch, cfg.in_channels, lh, lw, generator=gen),
        encoder_hidden_states=torch.randn(batch, text_len, cfg.joint_attention_dim, generator=gen),
        pooled_projections=torch.randn(batch, cfg.pooled_projection_dim, generator=gen),
        condition_pooled_projections=torch.randn(batch, cfg.pooled_projection_dim, generator=gen),
        timestep=torch.full((batch,), float(timestep)),
        rts_uniform=torch.rand(batch * N, cfg.expert_nums, generator=gen),
    )
