"""CPU: mathematical properties of the oracle's restatements of THIRD-PARTY arithmetic (diffusers embeddings / RoPE / norms), which no
reference golden can pin ("parity unpinned"). They do not prove equality with diffusers; they rule out the classic
restatement slips (sin / cos order, pairing of the rotated halves, frequency exponent, float64 angles, eps placement)."""
import math

import torch

from oracle import unigen_oracle as O
from oracle import unigen_sd3_oracle as S


def test_rope_is_a_rotation_that_only_sees_relative_positions():
    axes = (8, 28, 28)
    dh = sum(axes)
    g = torch.Generator().manual_seed(0)
    q, k = torch.randn(1, 2, 5, dh, generator=g), torch.randn(1, 2, 5, dh, generator=g)
    ids = torch.tensor([[0, 0, 0], [0, 3, 1], [0, 7, 9], [0, 2, 2], [0, 11, 4]], dtype=torch.float32)
    rope = O.flux_pos_embed(ids, axes)
    assert rope[0].shape == rope[1].shape == (5, dh)
    qr, kr = O.apply_rotary_emb(q, rope), O.apply_rotary_emb(k, rope)
    torch.testing.assert_close(qr.norm(dim=-1), q.norm(dim=-1), rtol=1e-5, atol=1e-5)  # a rotation
    torch.testing.assert_close(qr[:, :, 0], q[:, :, 0])  # position 0: cos = 1, sin = 0
    # shifting every position by the same offset leaves all q . k products unchanged
    shifted = O.flux_pos_embed(ids + torch.tensor([0.0, 5.0, 17.0]), axes)
    qs, ks = O.apply_rotary_emb(q, shifted), O.apply_rotary_emb(k, shifted)
    torch.testing.assert_close(qs @ ks.transpose(-1, -2), qr @ kr.transpose(-1, -2), rtol=1e-4, atol=1e-4)
    # pairing: channels (2i, 2i+1) rotate together by angle pos * theta^(-2i/d) — check the first pair of axis 1 by hand
    pos, d = 3.0, axes[1]
    ang = pos * 1.0 / (10000.0 ** (0.0 / d))
    x = torch.zeros(1, 1, 1, dh)
    x[..., axes[0]] = 1.0  # first channel of axis 1
    y = O.apply_rotary_emb(x, O.flux_pos_embed(torch.tensor([[0.0, pos, 0.0]]), axes))
    assert abs(y[0, 0, 0, axes[0]].item() - math.cos(ang)) < 1e-6 and abs(y[0, 0, 0, axes[0] + 1].item() - math.sin(ang)) < 1e-6


def test_timestep_projection_is_cos_first_with_the_published_frequencies():
    t = torch.tensor([0.0, 1.0, 250.0])
    e = O.timesteps_proj(t, 256)
    assert e.shape == (3, 256)
    torch.testing.assert_close(e[0], torch.cat([torch.ones(128), torch.zeros(128)]))  # flip_sin_to_cos: [cos | sin]
    # frequency i = 10000^(-i / 128) (downscale_freq_shift = 0)
    assert abs(e[1, 128 + 64].item() - math.sin(10000.0 ** (-64 / 128))) < 1e-6
    assert abs(e[2, 1].item() - math.cos(250.0 * 10000.0 ** (-1 / 128))) < 1e-4


def test_norms_match_torch_and_keep_eps_inside_the_root():
    g = torch.Generator().manual_seed(1)
    x = torch.randn(3, 7, 64, generator=g) * 3 + 1
    w = torch.randn(64, generator=g)
    torch.testing.assert_close(O.rms_norm(x, w), torch.nn.functional.rms_norm(x, (64,), w, eps=1e-6))
    torch.testing.assert_close(O.layer_norm(x), (x - x.mean(-1, keepdim=True)) / torch.sqrt(x.var(-1, unbiased=False, keepdim=True) + 1e-6),
                               rtol=1e-5, atol=1e-5)
    assert torch.isfinite(O.rms_norm(torch.zeros(1, 64), w)).all()  # eps inside the root: zero rows stay finite


def test_sd3_sincos_table_and_crop():
    D, grid, base = 32, 6, 4
    t = S.sincos_pos_embed_2d(D, grid, base)
    assert t.shape == (1, grid * grid, D)
    tt = t.view(grid, grid, D)
    # first half of the channels encodes the x (column) coordinate, second half the y (row) coordinate; each [sin | cos]
    torch.testing.assert_close(tt[0, :, D // 2:], tt[0, :1, D // 2:].expand(grid, -1))  # row 0: the y half is constant along x
    torch.testing.assert_close(tt[:, 0, :D // 2], tt[:1, 0, :D // 2].expand(grid, -1))  # column 0: the x half is constant along y
    torch.testing.assert_close(tt[0, 0], torch.cat([torch.zeros(8), torch.ones(8), torch.zeros(8), torch.ones(8)]))  # origin: sin 0, cos 0
    # coordinates are scaled by base_size / grid_size: position 3 of 6 with base 4 has coordinate 2
    assert abs(tt[0, 3, 0].item() - math.sin(2.0)) < 1e-6
    c = S.cropped_pos_embed(t, grid, 2, 4)
    assert c.shape[-2] == 8 and torch.equal(c.reshape(2, 4, D), tt[2:4, 1:5])  # centre crop


def test_top1gating_capacity_and_aux_loss_formula():
    g = torch.Generator().manual_seed(2)
    S_, E = 40, 4
    logits = torch.randn(S_, E, generator=g)
    C = O.moe_capacity(S_, E)
    assert C == max(math.ceil(S_ / E), 4)
    l_aux, combine, dispatch, counts, sparse = O.top1gating(logits, C, torch.rand(S_, E, generator=g))
    gates = torch.softmax(logits, -1)
    mask = torch.nn.functional.one_hot(gates.argmax(-1), E).float()
    torch.testing.assert_close(l_aux, (gates.mean(0) * mask.mean(0)).sum() * E)  # DeepSpeed: mean(me * ce) * E * E
    assert torch.equal(counts, mask.sum(0).long()) and dispatch.sum(dim=(0, 2)).max() <= C
    kept = dispatch.any(dim=(1, 2))
    torch.testing.assert_close(combine.sum(dim=(1, 2))[kept], gates.max(-1).values[kept])  # a kept token carries its gate probability
