"""Multi-GPU partitioning of the denoiser forward on one 8 x B200 box (NVLink 5 / NVSwitch), one process per GPU.

* Batch data parallelism (BASELINE cfg3): samples are independent -> `dp_shard`; NO data-path collective.
* Ulysses sequence parallelism (cfg4 / long multi-condition sequences): the ONE exchange step of the path is attention.
  Every rank owns S/P rows of the joint [text | image] sequence for all GEMM / LayerNorm / RMSNorm / RoPE work
  (token-local); around attention an all-to-all turns token shards x all heads into all tokens x H/P heads and back
  (4 exchanges per attention: Q, K, V in, O out; 2*S/P*D*2 bytes * (P-1)/P per tensor per rank). QK-RMSNorm and RoPE
  are per token and per head, so they run BEFORE the exchange on the local rows.
  The reference has no sequence parallelism at all — this is new design; its contract is numerical
  equality with the single-GPU forward (tests/test_parallel_gpu.py, tests/test_parallel_cpu.py).
  The CoMoE pre-stage (3.6 % of the step, once per step) needs a global top-C token selection per expert; it is computed
  REPLICATED on every rank from an all-gather of the residual stream (28 MB), which keeps routing bit-identical to
  the single-GPU run.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Tuple

import torch
import torch.distributed as dist

from . import ops
from .model import UniGenFlux


def dp_shard(n_samples: int, world: int, rank: int) -> range:
    """Contiguous block of sample indices owned by `rank` (ragged tails go to the first ranks)."""
    base, rem = divmod(n_samples, world)
    start = rank * base + min(rank, rem)
    return range(start, start + base + (1 if rank < rem else 0))


def sp_row_split(T: int, N: int, world: int, rank: int) -> Tuple[int, int, int, int]:
    """Rows of the joint [text | image] sequence owned by `rank`: (row0, rows, local_text_rows, first_local_image_token)."""
    S = T + N
    if S % world:
        raise ValueError(f"joint sequence {S} is not divisible by the sequence-parallel world size {world}")
    rows = S // world
    row0 = rank * rows
    t_loc = min(max(T - row0, 0), rows)
    return row0, rows, t_loc, max(row0 - T, 0)


class UlyssesExchange:
    """The all-to-all pair around attention for ONE sample (B = 1).

    seq_to_heads: local fused QKV rows [S_loc, 3*D] -> (q, k, v) each [S, D/P] (all tokens, this rank's heads)
    heads_to_seq: attention output [S, D/P] -> local rows of the [S_loc, D] (all heads) output view
    `copy` is the strided row-copy primitive (ops.copy on the GPU; a torch fallback is injected by the CPU tests only)."""

    def __init__(self, group, world: int, s_loc: int, d: int, device, dtype, copy: Optional[Callable] = None):
        self.group, self.P, self.s_loc, self.d = group, world, s_loc, d
        self.hd = d // world
        if d % world:
            raise ValueError("hidden size must be divisible by the sequence-parallel world size (whole heads per rank)")
        P, hd = world, self.hd
        self.send = torch.empty(3 * P, s_loc, hd, device=device, dtype=dtype)
        self.recv = torch.empty(3, P * s_loc, hd, device=device, dtype=dtype)
        self.o_full = torch.empty(P * s_loc, hd, device=device, dtype=dtype)
        self.o_recv = torch.empty(P, s_loc, hd, device=device, dtype=dtype)
        self.copy = copy or ops.copy

    def seq_to_heads(self, qkv_local: torch.Tensor):
        """qkv_local: [S_loc, 3*D] view (any row stride). Column block j = which*P + p (which in q,k,v) -> send[j]."""
        P, hd, s_loc = self.P, self.hd, self.s_loc
        src = torch.as_strided(qkv_local, (3 * P, s_loc, hd), (hd, qkv_local.stride(0), 1), qkv_local.storage_offset())
        self.copy(src, self.send)
        for w in range(3):
            dist.all_to_all_single(self.recv[w].view(-1), self.send[w * P:(w + 1) * P].reshape(-1), group=self.group)
        return self.recv[0], self.recv[1], self.recv[2]

    def heads_to_seq(self, out_local: torch.Tensor):
        """o_full [S, hd] (token-major == [P, S_loc, hd]) -> out_local [S_loc, D] view (any row stride)."""
        P, hd, s_loc = self.P, self.hd, self.s_loc
        dist.all_to_all_single(self.o_recv.view(-1), self.o_full.view(-1), group=self.group)
        dst = torch.as_strided(out_local, (P, s_loc, hd), (hd, out_local.stride(0), 1), out_local.storage_offset())
        self.copy(self.o_recv, dst)
        return out_local


class SequenceParallelUniGenFlux(UniGenFlux):
    """UniGenFlux whose main blocks run on S/P rows per rank with Ulysses attention (module docstring). Same public API;
    every rank passes the FULL inputs and receives the FULL velocity (all-gathered)."""

    def __init__(self, arch=None, device="cuda", group=None, **config):
        super().__init__(arch, device, **config)
        if not dist.is_initialized():
            raise ops.UgError("SequenceParallelUniGenFlux needs an initialised torch.distributed process group")
        self.sp_group = group
        self.sp_world = dist.get_world_size(group)
        self.sp_rank = dist.get_rank(group)
        if self.arch.num_attention_heads % self.sp_world:
            raise ops.UgError(f"{self.arch.num_attention_heads} heads are not divisible by {self.sp_world} ranks")
        self._sp_active = False
        self._xchg = None

    def _attend(self, buf, S: int, out: torch.Tensor):
        if not self._sp_active:
            return super()._attend(buf, S, out)
        a = self.arch
        x = self._xchg
        q, k, v = x.seq_to_heads(buf.QKV[0, :S])
        ops.attention(q.unsqueeze(0), k.unsqueeze(0), v.unsqueeze(0), x.o_full.unsqueeze(0), a.num_attention_heads // self.sp_world,
                      a.attention_head_dim, variant=self.attn_variant)
        x.heads_to_seq(out[0])
        return out

    def _forward_impl(self, conditioning_scale, hs, es, pooled, timestep, guidance, txt_ids, img_ids, **cond):
        a = self.arch
        P, rank = self.sp_world, self.sp_rank
        n_cond = self.condition_nums
        cs = [ops.to_bf16(cond[f"cs{c}"].contiguous()) for c in range(n_cond)]
        cond_pooled = [cond[f"cp{c}"] for c in range(n_cond)]
        condition_ids = [cond[f"cid{c}"] for c in range(n_cond)]
        rts_uniform = [cond[f"u{c}"] for c in range(n_cond)]
        D = self.inner_dim
        B, N, _ = hs.shape
        T = es.shape[1]
        S = T + N
        if B != 1:
            raise ops.UgError("sequence parallelism shards ONE sample across ranks (use batch data-parallelism for B > 1)")
        row0, s_loc, t_loc, i0 = sp_row_split(T, N, P, rank)
        n_img = s_loc - t_loc
        buf = self._workspace(B, N, T)
        dev = self.device_
        if self._xchg is None or self._xchg.s_loc != s_loc:
            self._xchg = UlyssesExchange(self.sp_group, P, s_loc, D, dev, torch.bfloat16)
            self._sp_buf = dict(XL=torch.empty(1, s_loc, D, device=dev, dtype=torch.bfloat16),
                                NOL=torch.empty(1, s_loc, D, device=dev, dtype=torch.bfloat16),
                                OUTL=torch.empty(1, s_loc, a.in_channels, device=dev, dtype=torch.bfloat16),
                                OUTF=torch.empty(1, S, a.in_channels, device=dev, dtype=torch.bfloat16))
        XL = self._sp_buf["XL"]
        xl_txt, xl_img = XL[:, :t_loc], XL[:, t_loc:]
        hs, es = ops.to_bf16(hs.contiguous()), ops.to_bf16(es.contiguous())
        gv = self.gemm_variant

        # ---- embeddings on the local rows only ----
        if n_img:
            ops.gemm(hs[:, i0:i0 + n_img], self.x_embedder_w[0], out=xl_img, bias=self.x_embedder_w[1], variant=gv)
        if t_loc:
            ops.gemm(es[:, row0:row0 + t_loc], self.context_embedder_w[0], out=xl_txt, bias=self.context_embedder_w[1], variant=gv)
        t_emb = ops.timestep_embedding(timestep * 1000.0)
        g_emb = ops.timestep_embedding(guidance * 1000.0) if guidance is not None else None
        self._time_text(self.time_text, t_emb, pooled, buf.temb, buf.tmp, g_emb)
        ctrl_pooled = pooled if self.use_pooled_prompt_embeds else torch.zeros_like(pooled)
        self._time_text(self.control_time_text, t_emb, ctrl_pooled, buf.ctemb, buf.tmp, g_emb)
        for c in range(n_cond):
            self._time_text(self.control_condition, t_emb, cond_pooled[c], buf.cdtemb_c[c], buf.tmp, g_emb)
            self._time_text(self.control_condition, t_emb, cond_pooled[c], buf.cdtemb, buf.tmp, g_emb, accumulate=c > 0)
        ops.rope_table(torch.cat([txt_ids, img_ids], 0), a.axes_dims_rope, a.theta, out=buf.rope)
        rope_loc = buf.rope[row0:row0 + s_loc]

        # ---- AdaLN vectors of every block (replicated: step constants) ----
        slot = 0
        m_double, m_cdouble, m_single, m_csingle, mods_s0 = [], [], [], [], []
        for w in self.double:
            m_double.append((self._mods(buf, slot, 6, w.norm1, buf.temb), self._mods(buf, slot + 6, 6, w.norm1_ctx, buf.temb)))
            slot += 12
        for w in self.ctrl_double:
            m_cdouble.append((self._mods(buf, slot, 6, w.norm1, buf.cdtemb), self._mods(buf, slot + 6, 6, w.norm1_ctx, buf.cdtemb)))
            slot += 12
        for w in self.single:
            m_single.append(self._mods(buf, slot, 3, w.norm, buf.temb)); slot += 3
        for w in self.ctrl_single:
            m_csingle.append(self._mods(buf, slot, 3, w.norm, buf.cdtemb)); slot += 3
        for c in range(n_cond):
            mods_s0.append((self._mods(buf, slot, 6, self.shared[0].norm1, buf.cdtemb_c[c]),
                            self._mods(buf, slot + 6, 6, self.shared[0].norm1_ctx, buf.cdtemb_c[c])))
            slot += 12
        mods_s1 = (self._mods(buf, slot, 6, self.shared[1].norm1, buf.ctemb),
                   self._mods(buf, slot + 6, 6, self.shared[1].norm1_ctx, buf.ctemb))
        slot += 12
        m_out = self._mods(buf, slot, 2, self.norm_out_w, buf.temb)

        # ---- double blocks on token shards ----
        route = None
        n_cd = len(self.ctrl_double)
        cenc_loc = None
        self._sp_active = True
        try:
            for i, w in enumerate(self.double):
                self._double_block(buf, w, m_double[i][0], m_double[i][1], xl_img, xl_txt, xl_img, xl_txt, rope_loc)
                j = int(i / (len(self.double) / n_cd))
                if route is None:
                    # CoMoE pre-stage, replicated: gather the residual stream once, route / run experts on every rank
                    self._sp_active = False
                    dist.all_gather_into_tensor(buf.X.view(-1), XL.view(-1), group=self.sp_group)
                    x_txt, x_img = buf.X[:, :T], buf.X[:, T:]
                    ops.gemm(x_txt, self.control_context_embedder_w[0], out=buf.CENC, bias=self.control_context_embedder_w[1], variant=gv)
                    for c in range(n_cond):
                        route = self._prestage(buf, B, N, T, x_img, cs[c], pooled, cond_pooled[c], rts_uniform[c], mods_s0[c],
                                               mods_s1, c, txt_ids, img_ids, condition_ids[c])
                    self._sp_active = True
                    ctrl_in = buf.CIN[:, i0:i0 + n_img]
                    cenc_loc = buf.CENC[:, row0:row0 + t_loc]
                else:
                    ctrl_in = xl_img
                ch = buf.CH[:, :n_img]
                self._double_block(buf, self.ctrl_double[j], m_cdouble[j][0], m_cdouble[j][1], ctrl_in, cenc_loc, ch, None, rope_loc)
                if n_img:
                    wa = self.add_double[j]
                    ops.gemm(ch, wa[0], out=xl_img, bias=wa[1], alpha=float(conditioning_scale), residual=xl_img, variant=gv)
            # ---- single blocks ----
            n_cs = len(self.ctrl_single)
            csl = buf.CS[:, :s_loc]
            for i, w in enumerate(self.single):
                self._single_block(buf, w, m_single[i], XL, XL, rope_loc)
                if n_cs:
                    j = int(i / (len(self.single) / n_cs))
                    self._single_block(buf, self.ctrl_single[j], m_csingle[j], XL, csl, rope_loc)
                    wa = self.add_single[j]
                    if self.single_block_control_method == "overall_add":
                        ops.gemm(csl, wa[0], out=XL, bias=wa[1], alpha=float(conditioning_scale), residual=XL, variant=gv)
                    elif n_img:
                        ops.gemm(csl[:, t_loc:], wa[0], out=xl_img, bias=wa[1], alpha=float(conditioning_scale), residual=xl_img, variant=gv)
        finally:
            self._sp_active = False
        # ---- norm_out + proj_out on the local rows, then gather the velocity ----
        sb = self._sp_buf
        ops.ln_modulate(XL, sb["NOL"], m_out[1], m_out[0])
        ops.gemm(sb["NOL"], self.proj_out_w[0], out=sb["OUTL"], bias=self.proj_out_w[1], variant=gv)
        dist.all_gather_into_tensor(sb["OUTF"].view(-1), sb["OUTL"].view(-1), group=self.sp_group)
        self._last_route = route
        return sb["OUTF"][:, T:], dict(moe_loss=route["l_aux"][0] * 0.1), dict(expert_counts=route["exp_counts"])
