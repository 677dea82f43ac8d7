"""Golden vectors for the SD3.5 path from the REAL reference code in /root/reference.

Same method as make_golden.py (stub modules for the absent third-party imports, then drive the reference-OWNED
functions with real tensors).  Pinned here:

  * `adanorm_forward`, `sd35adanormX_forward`, `adanormContinuous_forward`     src/UniGenUtils.py:340-373
    (per-sample 2-D and per-token 3-D `emb`)
  * `JointTransformerBlock.forward`                                            src/UniGenUtils.py:440-522
    (plain / dual attention / context_pre_only / per-token temb), `SD3SingleTransformerBlock.forward` :386-414
  * `MOELayer.forward` + `UniGenBase.expert_forward` (transformer-block experts, per-token temb) + `moe_forward`
    with the shared experts                                                    src/UniGenTransformer.py:222-296
  * `UniGenSD3.preprocess_moe_forward` / `control_forward` / `base_forward`   src/UniGenTransformer.py:498-623
    with affine stand-in blocks (weave order, first-call substitution, which tensors reach the MoE)
  * `UniGenSD3.forward` embedding order + un-patchify                         src/UniGenTransformer.py:625-710
  * `use_modulate=True`: `UniGenBase.expert_forward` modulated-linear branch (:252-255, real `modulated_flatten`)
    through the real `MOELayer`, `moe_forward` with `use_shared_expert` True / False (:279)

The attention / feed-forward / gate sub-modules those functions call are third-party (diffusers `Attention` +
`JointAttnProcessor2_0`, `FeedForward`, deepspeed `top1gating`): stand-ins written from the published algorithm are
plugged in, so those stay "parity unpinned".
Usage:  python tests/golden/make_golden_sd3.py   -> tests/golden/reference_golden_sd3.pt
"""
from __future__ import annotations

import functools
import sys
import types
from pathlib import Path

import torch
import torch.nn.functional as F
from torch import nn

sys.path.insert(0, str(Path(__file__).resolve().parent))
sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
from make_golden import import_reference  # noqa: E402

OUT = Path(__file__).resolve().parent / "reference_golden_sd3.pt"


class StandInAttention(nn.Module):
    """diffusers Attention(..., qk_norm='rms_norm' | None, added_kv_proj_dim=dim | None) + JointAttnProcessor2_0, restated."""

    def __init__(self, D, H, added, context_pre_only, qk_norm):
        super().__init__()
        self.heads, self.context_pre_only = H, context_pre_only
        dh = D // H
        self.to_q, self.to_k, self.to_v = nn.Linear(D, D), nn.Linear(D, D), nn.Linear(D, D)
        self.to_out = nn.ModuleList([nn.Linear(D, D), nn.Identity()])
        self.norm_q = nn.RMSNorm(dh, eps=1e-6) if qk_norm else None
        self.norm_k = nn.RMSNorm(dh, eps=1e-6) if qk_norm else None
        if added:
            self.add_q_proj, self.add_k_proj, self.add_v_proj = nn.Linear(D, D), nn.Linear(D, D), nn.Linear(D, D)
            self.to_add_out = nn.Linear(D, D) if not context_pre_only else None
            self.norm_added_q = nn.RMSNorm(dh, eps=1e-6) if qk_norm else None
            self.norm_added_k = nn.RMSNorm(dh, eps=1e-6) if qk_norm else None

    def forward(self, hidden_states, encoder_hidden_states=None, **kw):
        B, H = hidden_states.shape[0], self.heads
        hd = lambda t: t.view(B, -1, H, t.shape[-1] // H).transpose(1, 2)  # noqa: E731
        q, k, v = hd(self.to_q(hidden_states)), hd(self.to_k(hidden_states)), hd(self.to_v(hidden_states))
        if self.norm_q is not None:
            q, k = self.norm_q(q), self.norm_k(k)
        if encoder_hidden_states is not None:
            cq, ck, cv = (hd(self.add_q_proj(encoder_hidden_states)), hd(self.add_k_proj(encoder_hidden_states)),
                          hd(self.add_v_proj(encoder_hidden_states)))
            if self.norm_added_q is not None:
                cq, ck = self.norm_added_q(cq), self.norm_added_k(ck)
            q, k, v = torch.cat([q, cq], 2), torch.cat([k, ck], 2), torch.cat([v, cv], 2)
        o = F.scaled_dot_product_attention(q, k, v, dropout_p=0.0, is_causal=False)
        o = o.transpose(1, 2).reshape(B, -1, H * q.shape[-1])
        if encoder_hidden_states is not None:
            n = hidden_states.shape[1]
            o, co = o[:, :n], o[:, n:]
            if not self.context_pre_only:
                co = self.to_add_out(co)
            return self.to_out[1](self.to_out[0](o)), co
        return self.to_out[1](self.to_out[0](o))


class StandInFF(nn.Module):
    """diffusers FeedForward(dim, dim, activation_fn='gelu-approximate') with its parameter names (net.0.proj, net.2)."""

    def __init__(self, D):
        super().__init__()
        proj = nn.Module()
        proj.proj = nn.Linear(D, 4 * D)
        self.net = nn.ModuleList([proj, nn.Identity(), nn.Linear(4 * D, D)])

    def forward(self, x):
        return self.net[2](F.gelu(self.net[0].proj(x), approximate="tanh"))


class _Ada(nn.Module):
    def __init__(self, D, chunks, fwd):
        super().__init__()
        self.emb, self.silu, self.linear = None, nn.SiLU(), nn.Linear(D, chunks * D)
        self.norm = nn.LayerNorm(D, elementwise_affine=False, eps=1e-6)
        self._fwd = fwd

    def forward(self, **kw):
        return self._fwd(self, **kw)


def randomize(mod: nn.Module, g, scale=0.3):
    with torch.no_grad():
        for n, p in mod.named_parameters():
            if p.dim() == 1 and ("norm_" in n):
                p.copy_(1 + 0.1 * torch.randn(p.shape, generator=g))
            else:
                p.copy_(torch.randn(p.shape, generator=g) * (scale / max(1.0, (p.shape[-1] / 16) ** 0.5)))
    return mod


def sd_of(mod: nn.Module):
    return {k: v.detach().clone() for k, v in mod.state_dict().items()}


def main():
    mods = import_reference()
    U, T = mods["src.UniGenUtils"], mods["src.UniGenTransformer"]
    from oracle.unigen_oracle import top1gating  # third-party gate restated (unpinned), used as the stand-in gate

    g = torch.Generator().manual_seed(3535)
    gold = {}
    D, H = 32, 4

    # 1. AdaLN forwards -------------------------------------------------------------------------------------------
    cases = []
    for kind, chunks, fn in (("zero", 6, U.adanorm_forward), ("zero_x", 9, U.sd35adanormX_forward),
                             ("continuous", 2, U.adanormContinuous_forward)):
        m = randomize(_Ada(D, chunks, None), g)
        x = torch.randn(2, 7, D, generator=g)
        for emb in (torch.randn(2, D, generator=g), torch.randn(2, 7, D, generator=g)):
            if kind == "continuous" and emb.dim() == 3:
                continue  # chunk(dim=1) of a per-token emb is meaningless; the reference never does it
            with torch.no_grad():
                out = fn(m, x, emb=emb) if kind == "zero" else fn(m, x, emb)
            out = out if isinstance(out, tuple) else (out,)
            cases.append(dict(kind=kind, x=x, emb=emb, w=sd_of(m), out=[o.detach() for o in out]))
    gold["adaln"] = cases

    # 2. block forwards -------------------------------------------------------------------------------------------
    def make_joint(dual, cpo, qk=True):
        blk = nn.Module()
        blk.use_dual_attention, blk.context_pre_only = dual, cpo
        blk.norm1 = _Ada(D, 9 if dual else 6, functools.partial(
            lambda module, hidden_states, emb, f: f(module, hidden_states, emb) if dual else f(module, hidden_states, emb=emb),
            f=U.sd35adanormX_forward if dual else U.adanorm_forward))
        blk.norm1_context = _Ada(D, 2 if cpo else 6, functools.partial(
            lambda module, hidden_states, emb, f: f(module, hidden_states, emb) if cpo else f(module, hidden_states, emb=emb),
            f=U.adanormContinuous_forward if cpo else U.adanorm_forward))
        blk.attn = StandInAttention(D, H, True, cpo, qk)
        blk.attn2 = StandInAttention(D, H, False, None, qk) if dual else None
        blk.norm2 = nn.LayerNorm(D, elementwise_affine=False, eps=1e-6)
        blk.ff = StandInFF(D)
        if not cpo:
            blk.norm2_context = nn.LayerNorm(D, elementwise_affine=False, eps=1e-6)
            blk.ff_context = StandInFF(D)
        blk._chunk_size, blk._chunk_dim = None, 0
        blk.forward = lambda *a, **k: U.JointTransformerBlock.forward(blk, *a, **k)
        return randomize(blk, g)

    def make_single():
        blk = nn.Module()
        blk.norm1 = _Ada(D, 6, lambda module, hidden_states, emb: U.adanorm_forward(module, hidden_states, emb=emb))
        blk.attn = StandInAttention(D, H, False, None, False)
        blk.norm2 = nn.LayerNorm(D, elementwise_affine=False, eps=1e-6)
        blk.ff = StandInFF(D)
        blk.forward = lambda *a, **k: U.SD3SingleTransformerBlock.forward(blk, *a, **k)
        return randomize(blk, g)

    blocks = []
    for dual, cpo, per_token in ((False, False, False), (True, False, False), (True, True, False), (False, False, True)):
        blk = make_joint(dual, cpo)
        n, t = 6, (6 if per_token else 5)
        h, c = torch.randn(2, n, D, generator=g), torch.randn(2, t, D, generator=g)
        temb = torch.randn(2, n, D, generator=g) if per_token else torch.randn(2, D, generator=g)
        with torch.no_grad():
            enc_out, h_out = blk.forward(h, c, temb)
        blocks.append(dict(kind="joint", dual=dual, cpo=cpo, h=h, c=c, temb=temb, w=sd_of(blk),
                           enc_out=None if enc_out is None else enc_out.detach(), h_out=h_out.detach()))
    for per_token in (False, True):
        blk = make_single()
        x = torch.randn(1, 9, D, generator=g)
        temb = torch.randn(1, 9, D, generator=g) if per_token else torch.randn(1, D, generator=g)
        with torch.no_grad():
            y = blk.forward(x, temb)
        blocks.append(dict(kind="single", x=x, temb=temb, w=sd_of(blk), y=y.detach()))
    gold["blocks"] = blocks
    gold["heads"] = H

    # 3. MOELayer + expert_forward (transformer-block experts) + moe_forward with shared experts ------------------
    B, N, Tn, E, P = 2, 12, 5, 3, 8
    hidden, cond = torch.randn(B, N, D, generator=g), torch.randn(B, N, D, generator=g)
    enc = torch.randn(B, Tn, D, generator=g)
    temb, ctemb = torch.randn(B, D, generator=g), torch.randn(B, D, generator=g)
    pooled, cpooled = torch.randn(B, P, generator=g), torch.randn(B, P, generator=g)
    rts = torch.rand(B * N, E, generator=g)
    wg = torch.randn(E, D, generator=g) * 0.5
    experts = nn.ModuleList([nn.ModuleList([make_single(), make_single()]) for _ in range(E)])
    shared = [make_joint(False, False), make_joint(True, True)]
    C = max(-(-B * N // E), 4)

    class Gate(nn.Module):
        def forward(self, reshaped_input, used_token=None):
            l_aux, combine, dispatch, counts, _ = top1gating(F.linear(reshaped_input.float(), wg), C, rts)
            return l_aux, combine, dispatch, counts

    fake = types.SimpleNamespace(num_local_experts=E, use_modulate=False, use_rope=False, use_shared_expert=True,
                                 shared_expert=[shared[0].forward, shared[1].forward])

    class ExpertsFn(nn.Module):  # stands for deepspeed Experts whose .forward the reference re-binds (:174)
        def forward(self, **kw):
            return T.UniGenBase.expert_forward(fake, **kw)

    layer = U.MOELayer(Gate(), ExpertsFn(), "ep_size_1", 1, E)
    layer.experts.deepspeed_experts = experts
    fake.moe = types.SimpleNamespace(moe_layer=layer)
    with torch.no_grad():
        (eh, ec), l_aux, counts = T.UniGenBase.moe_forward(
            fake, hidden_states=hidden, condition_hidden_states=cond, encoder_hidden_states=enc, temb=temb,
            condition_temb=ctemb, condition_pooled_projections=cpooled, pooled_projections=pooled,
            joint_attention_kwargs=dict())
    gold["moe"] = dict(hidden=hidden, cond=cond, enc=enc, temb=temb, ctemb=ctemb, pooled=pooled, cpooled=cpooled, rts=rts,
                       wg=wg, E=E, C=C, experts=[[sd_of(b) for b in pair] for pair in experts],
                       shared=[sd_of(s) for s in shared], out_hidden=eh.detach(), out_cond=ec.detach(),
                       l_aux=l_aux.detach(), counts=counts.detach())

    # 4. UniGenSD3 weave: preprocess_moe_forward / control_forward / base_forward with affine stand-ins ------------
    def run_weave(n_base, n_ctrl):
        Dm, Tm, Nm = 4, 3, 5
        gg = torch.Generator().manual_seed(11 + n_base * 10 + n_ctrl)
        calls = []

        def mk(tag, i, last=False):
            a, b = 1.0 + 0.01 * (i + 1), 0.1 * (i + 1)

            def blk(hidden_states, encoder_hidden_states, temb, joint_attention_kwargs=None):
                calls.append((tag, i))
                enc_o = None if last else encoder_hidden_states * a + temb[:, None] * 0.01
                return enc_o, hidden_states * a + b + encoder_hidden_states.mean(1, keepdim=True) * 0.05 + temb[:, None] * 0.02
            return blk

        adders = [nn.Linear(Dm, Dm) for _ in range(n_ctrl)]
        with torch.no_grad():
            for m_ in adders:
                for p_ in m_.parameters():
                    p_.copy_(torch.randn(p_.shape, generator=gg) * 0.2)
        seen = {}

        def moe(**kw):
            seen.update({k: (v.detach().clone() if isinstance(v, torch.Tensor) else v) for k, v in kw.items()})
            return ((kw["hidden_states"] * 0.5 + 0.1, kw["condition_hidden_states"] * 0.25 - 0.1), torch.tensor(0.5),
                    torch.tensor([3, 2]))

        fs = types.SimpleNamespace(
            transformer_blocks=[mk("base", i, last=(i == n_base - 1)) for i in range(n_base)],
            control_transformer_blocks=[mk("ctrl", i + 50) for i in range(n_ctrl)], controlnet_add_blocks=adders,
            use_rope=False, use_encoder_hidden_states=True, use_pooled_prompt_embeds=True, cn_method="add",
            control_pos_embed_input=lambda x: x.flatten(2).transpose(1, 2)[..., :Dm] * 2.0 + 0.3,
            control_time_text_embed=lambda t, p: t[:, None] * 0.001 + p[:, :Dm],
            control_condition_embed=lambda t, p: t[:, None] * 0.002 - p[:, :Dm],
            control_context_embedder=lambda e: e * 1.5 - 0.2, moe=moe)
        fs.preprocess_moe_forward = lambda *a, **k: T.UniGenSD3.preprocess_moe_forward(fs, *a, **k)
        fs.control_forward = lambda *a, **k: T.UniGenSD3.control_forward(fs, *a, **k)
        h0, e0 = torch.randn(1, Nm, Dm, generator=gg), torch.randn(1, Tm, Dm, generator=gg)
        cond_lat = torch.randn(1, Dm, Nm, 1, generator=gg)
        temb_ = torch.randn(1, Dm, generator=gg)
        pooled_, cpooled_ = torch.randn(1, Dm + 2, generator=gg), torch.randn(1, Dm + 2, generator=gg)
        ts = torch.tensor([417.0])
        with torch.no_grad():
            res = T.UniGenSD3.base_forward(fs, hidden_states=h0, condition_hidden_states=cond_lat, encoder_hidden_states=e0,
                                           pooled_projections=pooled_, condition_pooled_projections=cpooled_, timestep=ts,
                                           conditioning_scale=0.7, temb=temb_, joint_attention_kwargs=None,
                                           img_ids=None, prompt_ids=None, condition_ids=None)
        return dict(n_base=n_base, n_ctrl=n_ctrl, h0=h0, e0=e0, cond_lat=cond_lat, temb=temb_, pooled=pooled_,
                    cpooled=cpooled_, timestep=ts, adders=[sd_of(m_) for m_ in adders], calls=calls,
                    moe_kwargs={k: v for k, v in seen.items() if isinstance(v, torch.Tensor)},
                    out_hidden=res["blocks_hidden_states"].detach(), moe_loss=res["moe_loss"], exp_count=res["exp_count"])

    gold["weave"] = [run_weave(24, 24), run_weave(4, 4), run_weave(6, 3)]

    # 5. UniGenSD3.forward: embedding order + un-patchify ---------------------------------------------------------
    Bf, Cc, Hh, Ww, p = 2, 3, 4, 6, 2
    lat = torch.randn(Bf, Cc, Hh, Ww, generator=g)
    tok = torch.randn(Bf, (Hh // p) * (Ww // p), p * p * Cc, generator=g)
    fs = types.SimpleNamespace(
        use_rope=False, config=types.SimpleNamespace(patch_size=p), out_channels=Cc,
        pos_embed=lambda x: x.flatten(2).transpose(1, 2), time_text_embed=lambda t, pp: pp,
        context_embedder=lambda e: e, norm_out=lambda h, temb: h, proj_out=lambda h: tok,
        base_forward=lambda **k: dict(blocks_hidden_states=k["hidden_states"], moe_loss=torch.tensor(2.0), exp_count=torch.tensor([1])))
    with torch.no_grad():
        out, losses, outs = T.UniGenSD3.forward(fs, lat, condition_hidden_states=lat, encoder_hidden_states=torch.zeros(Bf, 2, 4),
                                                pooled_projections=torch.zeros(Bf, 4), condition_pooled_projections=torch.zeros(Bf, 4),
                                                timestep=torch.tensor([1.0, 1.0]))
    gold["unpatchify"] = dict(tokens=tok, h=Hh // p, w=Ww // p, p=p, c=Cc, out=out.detach(), moe_loss=losses["moe_loss"])

    # 6. use_modulate=True: modulated-linear experts through the real MOELayer + UniGenBase.expert_forward (:252-255) and
    #    moe_forward with and without the shared experts (:279) ---------------------------------------------------------
    g6 = torch.Generator().manual_seed(3636)
    mod_cases = []
    for use_shared in (True, False):
        hidden6, cond6 = torch.randn(B, N, D, generator=g6), torch.randn(B, N, D, generator=g6)
        enc6 = torch.randn(B, Tn, D, generator=g6)
        temb6, ctemb6 = torch.randn(B, D, generator=g6), torch.randn(B, D, generator=g6)
        pooled6, cpooled6 = torch.randn(B, P, generator=g6), torch.randn(B, P, generator=g6)
        rts6 = torch.rand(B * N, E, generator=g6)
        wg6 = torch.randn(E, D, generator=g6) * 0.5
        experts6 = nn.ModuleList([nn.ModuleList([randomize(nn.ModuleList([nn.Linear(D, D), nn.Linear(P, D)]), g6)
                                                 for _ in range(2)]) for _ in range(E)])
        shared6 = [make_joint(False, False), make_joint(True, True)]  # consumes g (after every earlier section)

        class Gate6(nn.Module):
            def forward(self, reshaped_input, used_token=None, wg_=wg6, rts_=rts6):
                l_aux, combine, dispatch, counts, _ = top1gating(F.linear(reshaped_input.float(), wg_), C, rts_)
                return l_aux, combine, dispatch, counts

        fake6 = types.SimpleNamespace(num_local_experts=E, use_modulate=True, use_rope=False, use_shared_expert=use_shared,
                                      shared_expert=[shared6[0].forward, shared6[1].forward])

        class ExpertsFn6(nn.Module):
            def forward(self, fake_=fake6, **kw):
                return T.UniGenBase.expert_forward(fake_, **kw)

        layer6 = U.MOELayer(Gate6(), ExpertsFn6(), "ep_size_1", 1, E)
        layer6.experts.deepspeed_experts = experts6
        fake6.moe = types.SimpleNamespace(moe_layer=layer6)
        with torch.no_grad():
            (eh6, ec6), l_aux6, counts6 = T.UniGenBase.moe_forward(
                fake6, hidden_states=hidden6, condition_hidden_states=cond6, encoder_hidden_states=enc6, temb=temb6,
                condition_temb=ctemb6, condition_pooled_projections=cpooled6, pooled_projections=pooled6,
                joint_attention_kwargs=dict())
        mod_cases.append(dict(use_shared_expert=use_shared, hidden=hidden6, cond=cond6, enc=enc6, temb=temb6, ctemb=ctemb6,
                              pooled=pooled6, cpooled=cpooled6, rts=rts6, wg=wg6, E=E, C=C, P=P,
                              experts=[[sd_of(b) for b in pair] for pair in experts6], shared=[sd_of(s_) for s_ in shared6],
                              out_hidden=eh6.detach(), out_cond=ec6.detach(), l_aux=l_aux6.detach(), counts=counts6.detach()))
    gold["moe_modulate"] = mod_cases

    torch.save(gold, OUT)
    print("wrote", OUT, {k: (len(v) if isinstance(v, (list, dict)) else type(v)) for k, v in gold.items()})


if __name__ == "__main__":
    main()
