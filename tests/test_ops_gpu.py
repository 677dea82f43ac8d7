"""GPU parity of every C-ABI op against the oracle / plain torch fp32 on the same (bf16-rounded) inputs.
Tolerances: bf16 storage with fp32 accumulation -> rel-L2 <= 6e-3 per op (stated per test); integer / index
outputs (routing, masks) bit-exact."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def rel_l2(got, want):
    got, want = got.float(), want.float()
    return ((got - want).norm() / want.norm().clamp_min(1e-12)).item()


def rnd(*shape, scale=1.0, dev="cuda"):
    return (torch.randn(*shape, device=dev) * scale).to(torch.bfloat16)


@pytest.fixture(autouse=True)
def _seed():
    torch.manual_seed(0)


@pytest.mark.parametrize("variant", [1, 2, 3, 4, 5, 6])
@pytest.mark.parametrize("shape", [(1, 128, 256, 64), (1, 200, 384, 192), (1, 300, 64, 384), (2, 1000, 1152, 384),
                                   (1, 768, 384, 1920)])
def test_gemm_plain(ug, variant, shape):
    B, R, N, K = shape
    a, w = rnd(B, R, K), rnd(N, K, scale=K ** -0.5)
    out = ug.gemm(a, w, variant=variant)
    want = a.float() @ w.float().t()
    assert rel_l2(out, want) < 6e-3


# the GEMM shapes of the BASELINE configs (SURVEY.md §7 minimum slice, (M, K, N)): q|k|v-sized, proj_mlp / ff1, proj_out of the
# single blocks (K = 5D), ff2; the last one has A = 104 MB > 40 MB, which takes the banded tile walk (ug_gemm.cu group_m)
REAL_SHAPES = [(4608, 3072, 3072), (4608, 3072, 12288), (4608, 15360, 3072), (4096, 12288, 3072), (16896, 3072, 3072)]


@pytest.mark.parametrize("variant", [0, 1, 2, 3, 4, 5, 6])
@pytest.mark.parametrize("mkn", REAL_SHAPES)
def test_gemm_real_shapes(ug, variant, mkn):
    """fp32 torch matmul (TF32 off) on the same bf16 inputs; bias epilogue; rel-L2 <= 6e-3 (bf16 out, fp32 accumulate)."""
    torch.backends.cuda.matmul.allow_tf32 = False
    M, K, N = mkn
    a, w, bias = rnd(1, M, K), rnd(N, K, scale=K ** -0.5), rnd(N)
    out = ug.gemm(a, w, bias=bias, variant=variant)
    want = a[0].float() @ w.float().t() + bias.float()
    assert rel_l2(out[0], want) < 6e-3
    # no row / column was skipped or written twice by the (banded) persistent tile walk: every element is close, not just the norm
    assert (out[0].float() - want).abs().max() < 0.15


@pytest.mark.parametrize("variant", [1, 2, 3, 4, 5, 6])
def test_gemm_real_shape_fused_epilogue(ug, variant):
    """The gated-residual epilogue (`h + gate * (x W^T + b)`, in place) at the to_out / ff2 shape of cfg3."""
    torch.backends.cuda.matmul.allow_tf32 = False
    M, K, N = 4096, 12288, 3072
    a, w, bias = rnd(1, M, K), rnd(N, K, scale=K ** -0.5), rnd(N)
    gate = torch.randn(1, N, device="cuda")
    h = rnd(1, M, N)
    h0 = h.float().clone()
    ug.gemm(a, w, out=h, bias=bias, gate=gate, residual=h, variant=variant)
    want = h0 + gate[:, None] * (a.float() @ w.float().t() + bias.float())
    assert rel_l2(h, want) < 6e-3


@pytest.mark.parametrize("variant", [1, 2, 3, 4, 5, 6])
def test_gemm_fused_epilogue_strided(ug, variant):
    B, R, N, K = 3, 333, 640, 256
    a = rnd(B, R + 50, K + 64)[:, 50:, 64:]
    w, bias = rnd(N, K, scale=K ** -0.5), rnd(N)
    gate = torch.randn(B, N, device="cuda")
    res = rnd(B, R, N + 128)[:, :, 128:]
    out_full = torch.zeros(B, R + 7, N + 64, device="cuda", dtype=torch.bfloat16)
    out = out_full[:, 7:, 64:]
    ug.gemm(a, w, out=out, bias=bias, gate=gate, alpha=0.5, act=ug.UG_ACT_GELU_TANH, residual=res, variant=variant)
    y = torch.nn.functional.gelu(a.float() @ w.float().t() + bias.float(), approximate="tanh")
    want = y * gate[:, None, :] * 0.5 + res.float()
    assert rel_l2(out, want) < 6e-3
    assert out_full[:, :7].abs().max() == 0 and out_full[:, :, :64].abs().max() == 0  # no out-of-bounds writes


@pytest.mark.parametrize("variant", [0, 4, 5, 6])
def test_gemm_batched_weights_and_inplace_residual(ug, variant):
    """Stacked-expert GEMM (per-batch weights + bias) and the in-place gated residual; variants 4-6 = the smem-staged TMA-store
    epilogue (residual slab TMA-loaded into the staging buffer, result TMA-stored over it)."""
    E, C, D = 6, 43, 384
    a, w, b = rnd(E, C, D), rnd(E, D, D, scale=D ** -0.5), rnd(E, D)
    out = ug.gemm(a, w, bias=b, variant=variant)
    want = torch.einsum("ecd,end->ecn", a.float(), w.float()) + b.float()[:, None]
    assert rel_l2(out, want) < 6e-3
    h = rnd(2, 300, 384)
    h0 = h.clone()
    x, w2 = rnd(2, 300, 512), rnd(384, 512, scale=512 ** -0.5)
    gate = torch.randn(2, 384, device="cuda")
    ug.gemm(x, w2, out=h, gate=gate, residual=h, variant=variant)
    assert rel_l2(h, h0.float() + gate[:, None] * (x.float() @ w2.float().t())) < 6e-3


@pytest.mark.parametrize("n", [8, 40, 64, 72, 200])
def test_gemm_staged_epilogue_equals_direct_epilogue_bitwise(ug, n):
    """Same accumulators, same epilogue arithmetic: the staged (TMA-store) epilogue must reproduce the direct one bit for bit,
    also for output widths that end inside a 64-column staging slab / a 32-row lane quarter (hardware clipping)."""
    B, R, K = 2, 333, 320
    a, w, bias = rnd(B, R, K), rnd(n, K, scale=K ** -0.5), rnd(n)
    gate, res = torch.randn(B, n, device="cuda"), rnd(B, R, n)
    for direct, staged in ((1, 5), (2, 4), (3, 6)):
        ref = torch.full((B, R + 2, n + 8), 7.0, device="cuda", dtype=torch.bfloat16)
        got = ref.clone()
        kw = dict(bias=bias, gate=gate, alpha=0.7, act=ug.UG_ACT_GELU_TANH, residual=res)
        ug.gemm(a, w, out=ref[:, 1:-1, :n], variant=direct, **kw)
        ug.gemm(a, w, out=got[:, 1:-1, :n], variant=staged, **kw)
        assert torch.equal(got, ref)  # incl. the untouched guard rows / columns around the output view


def test_gemm_rejects_bad_arguments(ug):
    with pytest.raises(ug.UgError):
        ug.gemm(rnd(1, 16, 60), rnd(32, 60))  # K not a multiple of 8
    with pytest.raises(ug.UgError):
        ug.gemm(torch.zeros(1, 16, 64, dtype=torch.bfloat16), torch.zeros(32, 64, dtype=torch.bfloat16))  # CPU tensors


def _sdpa_ref(qkv, H, dh, mask=None):
    B, S = qkv.shape[:2]
    q, k, v = (qkv[:, :, i].float().reshape(B, S, H, dh).transpose(1, 2) for i in range(3))
    s = q @ k.transpose(-1, -2) / math.sqrt(dh)
    if mask is not None:
        s = s.masked_fill(~mask, float("-inf"))
    return (torch.softmax(s, -1) @ v).transpose(1, 2).reshape(B, S, H * dh)


@pytest.mark.parametrize("variant", [1, 2, 3, 4, 5, 6, 7, 9, 10])
@pytest.mark.parametrize("shape", [(1, 128, 1, 128), (1, 300, 3, 128), (2, 333, 2, 64), (1, 1024, 6, 64), (1, 1536, 4, 128)])
def test_attention_full(ug, variant, shape):
    B, S, H, dh = shape
    qkv = rnd(B, S, 3, H * dh)
    out = torch.zeros(B, S, H * dh, device="cuda", dtype=torch.bfloat16)
    ug.attention(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], out, H, dh, variant=variant)
    assert rel_l2(out, _sdpa_ref(qkv, H, dh)) < 6e-3


@pytest.mark.parametrize("variant", [1, 2, 3, 4, 5, 6, 7, 9, 10])
@pytest.mark.parametrize("strict", [False, True])
def test_attention_segment_mask_and_bit_exact_mask(ug, variant, strict):
    """[txt | img | c1 | c2] with the reference's visibility rule and the north-star's stricter one."""
    from oracle import unigen_oracle as O
    bounds = [0, 64, 264, 464, 600]
    vis = O.pvariant_visibility(2, strict=strict)
    S, H, dh = bounds[-1], 3, 128
    qkv = rnd(2, S, 3, H * dh)
    out = torch.zeros(2, S, H * dh, device="cuda", dtype=torch.bfloat16)
    ug.attention(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], out, H, dh, seg_bounds=bounds, seg_visible=vis, variant=variant)
    mask = O.segment_mask(bounds, vis)
    assert rel_l2(out, _sdpa_ref(qkv, H, dh, mask.cuda())) < 6e-3
    assert torch.equal(ug.expand_segment_mask(S, bounds, vis, "cuda").cpu(), mask)  # bit-exact mask construction


def test_ln_modulate(ug):
    from oracle import unigen_oracle as O
    for D in (384, 1536, 3072):
        x = rnd(2, 77, D, scale=2.0)
        shift, scale = torch.randn(2, D, device="cuda"), torch.randn(2, D, device="cuda")
        out = torch.empty_like(x)
        ug.ln_modulate(x, out, shift, scale)
        want = O.layer_norm(x.float()) * (1 + scale[:, None]) + shift[:, None]
        assert rel_l2(out, want) < 4e-3


def test_ln_modulate_persistent_kernel_many_rows(ug):
    """Enough row groups to take the persistent prefetching kernel (two blocks per SM walking the groups, next group's rows in
    flight), ragged last group, batch 2 and strided input rows."""
    from oracle import unigen_oracle as O
    D = 3072
    for B, R in ((1, 4608), (2, 1501)):
        xx = rnd(B, R, 2 * D, scale=2.0)
        x = xx[:, :, :D]  # row stride 2 D
        shift, scale = torch.randn(B, D, device="cuda"), torch.randn(B, D, device="cuda")
        out = torch.full((B, R + 1, D), 7.0, device="cuda", dtype=torch.bfloat16)
        ug.ln_modulate(x, out[:, :R], shift, scale)
        want = O.layer_norm(x.float()) * (1 + scale[:, None]) + shift[:, None]
        assert rel_l2(out[:, :R], want) < 4e-3
        assert torch.all(out[:, R] == 7.0)  # nothing written past the last row


@pytest.mark.parametrize("dh", [64, 128])
def test_qk_rmsnorm_rope(ug, dh):
    from oracle import unigen_oracle as O
    B, S, H = 2, 200, 3
    axes = (16, 56, 56) if dh == 128 else (8, 28, 28)
    ids = torch.stack([torch.zeros(S), torch.arange(S) // 16, torch.arange(S) % 16], 1).float()
    table = ug.rope_table(ids.cuda(), axes)
    cos, sin = O.flux_pos_embed(ids, axes)
    # table layout: (cos_i, sin_i) per pair
    assert torch.allclose(table.cpu().view(S, dh // 2, 2)[..., 0], cos[:, 0::2], atol=1e-6)
    assert torch.allclose(table.cpu().view(S, dh // 2, 2)[..., 1], sin[:, 0::2], atol=1e-6)
    qk = rnd(B, S, 2 * H * dh)
    w = (1 + 0.1 * torch.randn(2, dh)).to(torch.bfloat16).cuda()
    want = []
    for part in range(2):
        x = qk[:, :, part * H * dh:(part + 1) * H * dh].float().reshape(B, S, H, dh).transpose(1, 2).cpu()
        y = O.apply_rotary_emb(O.rms_norm(x, w[part].float().cpu()), (cos, sin))
        want.append(y.transpose(1, 2).reshape(B, S, H * dh))
    ug.qk_rmsnorm_rope(qk, 2 * H, dh, w, table, heads_per_weight=H)
    assert rel_l2(qk.cpu(), torch.cat(want, -1)) < 4e-3


def test_gemv_and_timestep_embedding(ug):
    from oracle import unigen_oracle as O
    for B in (1, 2, 5, 8):
        x = torch.randn(B, 768, device="cuda")
        w, b = rnd(1000, 768, scale=768 ** -0.5), rnd(1000)
        out = ug.gemv(x, w, b, silu_in=True)
        want = torch.nn.functional.silu(x) @ w.float().t() + b.float()
        assert rel_l2(out, want) < 1e-5
        out2 = ug.gemv(x, w, b, out=out.clone(), silu_out=True, accumulate=True)
        assert rel_l2(out2, want + torch.nn.functional.silu(x @ w.float().t() + b.float())) < 1e-5
    t = torch.tensor([1000.0, 250.0, 3.5], device="cuda")
    assert torch.allclose(ug.timestep_embedding(t).cpu(), O.timesteps_proj(t.cpu()), atol=2e-4)


def test_add_copy_cast(ug):
    a, b = rnd(2, 50, 384), rnd(2, 50, 384)
    big = torch.zeros(2, 80, 768, device="cuda", dtype=torch.bfloat16)
    ug.add(a, b, big[:, 30:, 384:])
    assert torch.equal(big[:, 30:, 384:], (a.float() + b.float()).to(torch.bfloat16))
    ug.copy(big[:, 30:, 384:], big[:, :50, :384])
    assert torch.equal(big[:, :50, :384], big[:, 30:, 384:])
    f = torch.randn(1000, device="cuda")
    assert torch.equal(ug.to_bf16(f), f.to(torch.bfloat16))
    assert torch.equal(ug.to_f32(a), a.float())


@pytest.mark.parametrize("tokens,experts,d", [(256, 6, 384), (1024, 6, 3072), (4096, 12, 512), (37, 6, 128), (4096, 6, 3072)])
def test_moe_route_bit_exact(ug, tokens, experts, d):
    """DeepSpeed top1gating + RTS: expert_idx / slot / slot_token / exp_counts bit-exact vs the oracle
    given the same uniform draw; prob and l_aux to fp32 rounding."""
    from oracle import unigen_oracle as O
    x = rnd(tokens, d)
    wg = ((torch.rand(experts, d, device="cuda") * 2 - 1) / math.sqrt(d)).float()
    wg[0] += 0.3 / math.sqrt(d)  # unbalance so that some expert overflows its capacity
    u = torch.rand(tokens, experts, device="cuda")
    C = O.moe_capacity(tokens, experts)
    r = ug.moe_route(x, wg, u, C)
    logits = (x.float() @ wg.t()).cpu()
    l_aux, combine, dispatch, counts, (idx, slot, prob) = O.top1gating(logits, C, u.cpu())
    # a near-tie in the fp32 logits may legitimately flip the argmax (different summation order): require none here
    top2 = logits.topk(2, dim=1).values
    assert (top2[:, 0] - top2[:, 1]).min() > 1e-5, "test seed produced a near-tie; pick another seed"
    assert torch.equal(r["expert_idx"].cpu().long(), idx)
    assert torch.equal(r["exp_counts"].cpu(), counts)
    assert torch.equal(r["slot"].cpu().long(), slot)
    st = torch.full((experts * C,), -1, dtype=torch.long)
    kept = slot >= 0
    st[idx[kept] * C + slot[kept]] = torch.arange(tokens)[kept]
    assert torch.equal(r["slot_token"].cpu().long(), st)
    assert torch.allclose(r["prob"].cpu(), prob, rtol=1e-5, atol=1e-6)
    assert abs(r["l_aux"].item() - l_aux.item()) < 1e-5
    assert (counts > C).any() or tokens < 64  # the dropping path was exercised


def test_moe_gather_experts_combine_matches_dense_einsum(ug):
    """Sparse gather -> modulate -> stacked expert GEMMs -> combine vs the reference's dense dispatch/combine algebra
    (oracle.moe_dispatch / expert math / moe_combine) on identical routing."""
    from oracle import unigen_oracle as O
    B, N, D, E, P = 2, 128, 384, 6, 768
    tokens = B * N
    C = O.moe_capacity(tokens, E)
    hid, cond = rnd(B, N, D), rnd(B, N, D)
    x = (hid.float() + cond.float()).to(torch.bfloat16).view(tokens, D)
    wg = ((torch.rand(E, D, device="cuda") * 2 - 1) / math.sqrt(D)).float()
    u = torch.rand(tokens, E, device="cuda")
    r = ug.moe_route(x, wg, u, C)
    Wc, bc, Wh, bh = rnd(E, D, D, scale=D ** -0.5), rnd(E, D), rnd(E, D, D, scale=D ** -0.5), rnd(E, D)
    sc, sh = torch.randn(B, E, D, device="cuda"), torch.randn(B, E, D, device="cuda")  # L^e(pooled[b])
    A = ug.moe_gather_modulate(cond.view(tokens, D), r["slot_token"], sc, E, C, N)
    Yc = ug.gemm(A.view(E, C, D), Wc, bias=bc)
    A2 = ug.moe_gather_modulate(hid.view(tokens, D), r["slot_token"], sh, E, C, N, addend=Yc.view(E * C, D))
    Yh = ug.gemm(A2.view(E, C, D), Wh, bias=bh)
    out_h, out_c = torch.empty(tokens, D, device="cuda", dtype=torch.bfloat16), torch.empty(tokens, D, device="cuda", dtype=torch.bfloat16)
    ug.moe_combine(Yh.view(E * C, D), r, C, out_h)
    ug.moe_combine(Yc.view(E * C, D), r, C, out_c)
    # dense reference algebra in fp32 with the oracle's masks
    logits = (x.float() @ wg.t()).cpu()
    _, combine, dispatch, _, _ = O.top1gating(logits, C, u.cpu())
    dh_, dc_ = O.moe_dispatch(dispatch, hid.float().cpu().view(tokens, D)), O.moe_dispatch(dispatch, cond.float().cpu().view(tokens, D))
    b_of = torch.arange(tokens) // N
    eh, ec = [], []
    for e in range(E):
        s_c = O.moe_dispatch(dispatch, sc[:, e].cpu()[b_of])[e]
        s_h = O.moe_dispatch(dispatch, sh[:, e].cpu()[b_of])[e]
        c2 = (dc_[e] * s_c) @ Wc[e].float().cpu().t() + bc[e].float().cpu()
        h2 = ((dh_[e] + c2) * s_h) @ Wh[e].float().cpu().t() + bh[e].float().cpu()
        eh.append(h2); ec.append(c2)
    want_h = O.moe_combine(combine, torch.stack(eh), hid.float().cpu().view(tokens, D))
    want_c = O.moe_combine(combine, torch.stack(ec), hid.float().cpu().view(tokens, D))
    assert rel_l2(out_c.cpu(), want_c) < 8e-3
    assert rel_l2(out_h.cpu(), want_h) < 1e-2
    dropped = (r["slot"] < 0)
    assert dropped.any() and out_h[dropped].abs().max() == 0  # dropped tokens are exact zeros


def test_ln_modulate_segs_equals_per_segment_calls(ug):
    B, D, bounds = 2, 384, [0, 33, 200, 201, 333]
    x = rnd(B, bounds[-1], D, scale=2.0)
    table = torch.randn(4, B, 6 * D, device="cuda")
    got = ug.ln_modulate_segs(x, torch.empty_like(x), table[0][:, 3 * D:4 * D], table[0][:, 4 * D:5 * D], bounds, B * 6 * D)
    want = torch.empty_like(x)
    for s in range(4):
        ug.ln_modulate(x[:, bounds[s]:bounds[s + 1]], want[:, bounds[s]:bounds[s + 1]], table[s][:, 3 * D:4 * D], table[s][:, 4 * D:5 * D])
    assert torch.equal(got, want)


@pytest.mark.parametrize("H,dh", [(24, 128), (5, 128), (24, 64), (9, 64)])
def test_qk_rmsnorm_rope_many_heads(ug, H, dh):
    """More heads than one warp pass covers (one warp per (row, head chunk)), incl. a ragged last chunk."""
    S = 77
    axes = (16, 56, 56) if dh == 128 else (8, 28, 28)
    ids = torch.stack([torch.zeros(S), torch.arange(S) // 9, torch.arange(S) % 9], 1).float().cuda()
    table = ug.rope_table(ids, axes)
    qk = rnd(1, S, 2 * H * dh)
    w = (1 + 0.1 * torch.randn(2, dh)).to(torch.bfloat16).cuda()
    x = qk.float().reshape(S, 2, H, dh)
    y = x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + 1e-6) * w.float()[None, :, None, :]
    cs = table.view(S, dh // 2, 2)
    y0, y1 = y[..., 0::2], y[..., 1::2]
    c, s_ = cs[:, None, None, :, 0], cs[:, None, None, :, 1]
    want = torch.stack([y0 * c - y1 * s_, y1 * c + y0 * s_], -1).reshape(1, S, 2 * H * dh)
    ug.qk_rmsnorm_rope(qk, 2 * H, dh, w, table, heads_per_weight=H)
    assert rel_l2(qk, want) < 4e-3


def _pools(P, nbytes):
    from unigen_b200 import _lib
    bufs = [torch.zeros(nbytes, dtype=torch.uint8, device="cuda") for _ in range(P)]
    tables = []
    for r in range(P):
        t = _lib.PeerTable()
        t.world, t.rank = P, r
        for i in range(P):
            t.base[i] = bufs[i].data_ptr()
        tables.append(t)
    return bufs, tables


@pytest.mark.parametrize("P", [1, 2, 4, 8])
@pytest.mark.parametrize("segmented", [False, True])
def test_ulysses_peer_exchange_emulated_ranks(ug, P, segmented):
    """The fused exchange kernels with P EMULATED ranks on one GPU (every "peer pool" is a local buffer, the ranks run one
    after the other): qkv_scatter + attention_peer with the uniform and the segment-sharded row maps reproduce the
    unsharded qk_rmsnorm_rope + (masked) attention bit for bit."""
    from oracle import unigen_oracle as O
    H, dh = 8, 64
    D, hd = H * dh, H * dh // P
    gb = [0, 64, 320, 448, 576] if segmented else [0, 576]      # global segment bounds, every bound a multiple of 8
    vis = O.pvariant_visibility(2) if segmented else None
    S = gb[-1]
    qkv = rnd(S, 3 * D)
    w = (1 + 0.1 * torch.randn(2, dh)).to(torch.bfloat16).cuda()
    ids = torch.stack([torch.zeros(S), torch.arange(S) // 24, torch.arange(S) % 24], 1).float().cuda()
    rope = ug.rope_table(ids, (8, 28, 28))
    # unsharded reference with the same kernels (variant pinned: the auto choice depends on the head count)
    ref_qkv = qkv.clone().unsqueeze(0)
    ug.qk_rmsnorm_rope(ref_qkv[:, :, :2 * D], 2 * H, dh, w, rope, heads_per_weight=H)
    want = torch.zeros(1, S, D, device="cuda", dtype=torch.bfloat16)
    ug.attention(ref_qkv[:, :, :D], ref_qkv[:, :, D:2 * D], ref_qkv[:, :, 2 * D:], want, H, dh, seg_bounds=gb if segmented else None,
                 seg_visible=vis, variant=1)
    off_recv, off_ao = 4096, 4096 + 3 * S * hd * 2
    bufs, tables = _pools(P, off_ao + (S // P) * D * 2 + 4096)
    # local rows of rank r: its shard of every segment, in segment order
    local = [[(gb[s] + r * (gb[s + 1] - gb[s]) // P, (gb[s + 1] - gb[s]) // P) for s in range(len(gb) - 1)] for r in range(P)]
    for r in range(P):
        for row0, n in local[r]:
            ug.qkv_scatter(tables[r], qkv[row0:row0 + n], H, dh, w, rope[row0:row0 + n].contiguous(), off_recv, S, row0)
    for r in range(P):
        recv = bufs[r][off_recv:off_recv + 3 * S * hd * 2].view(torch.bfloat16).view(3, 1, S, hd)
        ug.attention_peer(tables[r], recv[0], recv[1], recv[2], H // P, dh, off_ao, D, 0 if segmented else S // P,
                          seg_bounds=gb if segmented else None, seg_visible=vis, variant=1)
    torch.cuda.synchronize()
    got = torch.empty_like(want)
    for r in range(P):
        ao = bufs[r][off_ao:off_ao + (S // P) * D * 2].view(torch.bfloat16).view(S // P, D)
        lo = 0
        for row0, n in local[r]:
            got[0, row0:row0 + n] = ao[lo:lo + n]
            lo += n
    assert torch.equal(got, want)
    # broadcast gather + barrier (world 1 table: the flag round trip on the local pool)
    src = rnd(S // P, 64)
    for r in range(P):
        ug.peer_bcast_rows(tables[r], src, off_recv, 64, r * (S // P))
    full = bufs[P - 1][off_recv:off_recv + S * 64 * 2].view(torch.bfloat16).view(S, 64)
    assert all(torch.equal(full[r * (S // P):(r + 1) * (S // P)], src) for r in range(P))
    b1, t1 = _pools(1, 8192)
    for _ in range(3):
        ug.peer_barrier(t1[0])
    torch.cuda.synchronize()
    ctrl = b1[0][:264].view(torch.int32)
    assert ctrl[64].item() == 3 and ctrl[0].item() == 3 and ctrl[65].item() == 0  # epoch, own flag, no watchdog trip


@pytest.mark.parametrize("B", [1, 2, 3, 8])
def test_gemv_grouped_one_launch_equals_per_job_gemv(ug, B):
    """Every AdaLN linear of a step in ONE launch (device-resident job table): same numbers as one ug_gemv per job; sharding the
    group list over emulated ranks with a "peer pool" output reproduces the unsharded table on every rank."""
    from unigen_b200 import _lib
    torch.manual_seed(3)
    shapes = [(1152, 384), (2304, 384), (768, 768), (6144, 3072), (40, 384), (9216, 3072)]  # (n, k); 40: ragged last group
    xs = [torch.randn(B, 384, device="cuda"), torch.randn(B, 768, device="cuda"), torch.randn(B, 3072, device="cuda")]
    xof = {384: xs[0], 768: xs[1], 3072: xs[2]}
    total = sum(n for n, _ in shapes)
    out = torch.zeros(B, total, device="cuda")
    jobs, want, c0 = [], [], 0
    for j, (n, k) in enumerate(shapes):
        w, b = rnd(n, k, scale=k ** -0.5), (rnd(n) if j % 2 == 0 else None)
        silu_in = j % 3 != 1
        jobs.append((w, b, xof[k], out[:, c0:c0 + n], silu_in))
        want.append(ug.gemv(xof[k], w, b, silu_in=silu_in))
        ref = (torch.nn.functional.silu(xof[k]) if silu_in else xof[k]) @ w.float().t() + (b.float() if b is not None else 0)
        assert rel_l2(want[-1], ref) < 1e-5
        c0 += n
    plan = ug.GemvPlan(jobs, "cuda")
    assert plan.total_groups == sum((n + 3) // 4 for n, _ in shapes)
    ug.gemv_grouped(plan)
    assert rel_l2(out, torch.cat(want, 1)) < 1e-5  # fp32 summation order differs slightly from the per-job kernel
    # emulated sequence-parallel ranks: each computes its share of the groups and stores into EVERY "pool"
    P = 4
    pool_bytes = 4096 + B * total * 4
    pools = [torch.zeros(pool_bytes, dtype=torch.uint8, device="cuda") for _ in range(P)]
    for r in range(P):
        t = _lib.PeerTable()
        t.world, t.rank = P, r
        for i in range(P):
            t.base[i] = pools[i].data_ptr()

        import types as _types
        Pool = _types.SimpleNamespace(local=pools[r], nbytes=pool_bytes, table=t)  # what GemvPlan needs of a parallel.PeerPool

        view = pools[r][4096:].view(torch.float32).view(B, total)
        pj, c0 = [], 0
        for (w, b, x, _, s_), (n, k) in zip(jobs, shapes):
            pj.append((w, b, x, view[:, c0:c0 + n], s_))
            c0 += n
        ug.gemv_grouped(ug.GemvPlan(pj, "cuda", pool=Pool), rank=r, world=P)
    torch.cuda.synchronize()
    for r in range(P):
        assert torch.equal(pools[r][4096:].view(torch.float32).view(B, total), out)
    y = torch.empty_like(xs[2])
    assert torch.allclose(ug.silu(xs[2], y), torch.nn.functional.silu(xs[2]), rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("variant", [0, 1, 2, 4, 5])
@pytest.mark.parametrize("H,dh,rows", [(3, 128, 300), (6, 64, 333), (24, 128, 640)])
def test_gemm_fused_qk_rmsnorm_rope_epilogue(ug, variant, H, dh, rows):
    """north_star subsystem (1) "QK-RMSNorm and RoPE fused into the Q/K load": the q|k|v projection GEMM normalises every q / k
    head from the fp32 accumulator and rotates it in its epilogue (direct epilogue: variants 1 / 2; smem-staged TMA-store
    epilogue: 4 / 5, the default) == plain projection followed by the separate in-place pass, up to the bf16 rounding the
    separate pass applies BEFORE normalising. v columns are bit-identical to the plain projection."""
    from oracle import unigen_oracle as O
    D, K = H * dh, 256
    axes = (16, 56, 56) if dh == 128 else (8, 28, 28)
    x, w, bias = rnd(2, rows, K), rnd(3 * D, K, scale=K ** -0.5), rnd(3 * D)
    nw = (1 + 0.1 * torch.randn(2, dh)).to(torch.bfloat16).cuda()
    ids = torch.stack([torch.zeros(rows), torch.arange(rows) // 20, torch.arange(rows) % 20], 1).float().cuda()
    table = ug.rope_table(ids, axes)
    fused = ug.gemm(x, w, bias=bias, variant=variant, qk_norm=dict(weight=nw, head_dim=dh, d=D, cos_sin=table, eps=1e-6))
    plain = ug.gemm(x, w, bias=bias, variant=1)
    assert torch.equal(fused[:, :, 2 * D:], plain[:, :, 2 * D:])
    # fp32 reference straight from the definition (diffusers RMSNorm + apply_rotary_emb on the fp32 projection)
    y = (x.float() @ w.float().t() + bias.float())[:, :, :2 * D].reshape(2, rows, 2, H, dh)
    y = y * torch.rsqrt(y.pow(2).mean(-1, keepdim=True) + 1e-6) * nw.float()[None, None, :, None, :]
    cs = table.view(rows, dh // 2, 2)
    c, s_ = cs[None, :, None, None, :, 0], cs[None, :, None, None, :, 1]
    y0, y1 = y[..., 0::2], y[..., 1::2]
    want = torch.stack([y0 * c - y1 * s_, y1 * c + y0 * s_], -1).reshape(2, rows, 2 * D)
    assert rel_l2(fused[:, :, :2 * D], want) < 4e-3
    sep = plain.clone()
    ug.qk_rmsnorm_rope(sep[:, :, :2 * D], 2 * H, dh, nw, table, heads_per_weight=H)
    assert rel_l2(sep[:, :, :2 * D], want) < 6e-3  # the separate pass starts from bf16-rounded projections


@pytest.mark.parametrize("dh,H", [(128, 24), (64, 24)])
def test_attention_split_p_variant_is_bit_identical_at_full_size(ug, dh, H):
    """Variant 5 (P published in two 64-key halves, PV started on the first) issues the same MMAs on the same operands in the
    same order as variant 3: bit-identical output at the cfg3 / cfg5 sequence lengths."""
    S = 4608 if dh == 128 else 4429
    qkv = rnd(1, S, 3, H * dh)
    outs = []
    for variant in (3, 5, 7):  # 7 = the persistent form of 5 (one CTA per SM walking the (query tile, head) units)
        out = torch.zeros(1, S, H * dh, device="cuda", dtype=torch.bfloat16)
        ug.attention(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], out, H, dh, variant=variant)
        outs.append(out)
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])
    assert rel_l2(outs[1][:, :512], _sdpa_ref(qkv, H, dh)[:, :512]) < 6e-3


def test_attention_persistent_segment_mask_many_units(ug):
    """Persistent two-tile kernel with more work units than SMs AND a segment mask (units of different length, tile lists built
    one unit ahead): bit-identical to the one-unit-per-CTA kernel."""
    from oracle import unigen_oracle as O
    bounds = [0, 512, 2560, 3584, 4608]
    vis = O.pvariant_visibility(2)
    S, H, dh = bounds[-1], 12, 128
    qkv = rnd(2, S, 3, H * dh)
    outs = []
    for variant in (5, 7):
        out = torch.zeros(2, S, H * dh, device="cuda", dtype=torch.bfloat16)
        ug.attention(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], out, H, dh, seg_bounds=bounds, seg_visible=vis, variant=variant)
        outs.append(out)
    assert torch.equal(outs[0], outs[1])
    mask = O.segment_mask(bounds, vis)
    assert rel_l2(outs[1][:, ::7], _sdpa_ref(qkv, H, dh, mask.cuda())[:, ::7]) < 6e-3
