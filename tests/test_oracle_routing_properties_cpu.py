"""CPU: size-independent properties of the oracle's DeepSpeed top-1 gate restatement — the integer side of
the path that the CUDA routing kernels must reproduce bit for bit (tests/test_ops_gpu.py::test_moe_route_bit_exact)."""
import torch
from hypothesis import given, settings, strategies as st

from oracle import unigen_oracle as O


@settings(max_examples=40, deadline=None)
@given(tokens=st.integers(min_value=4, max_value=300), experts=st.integers(min_value=1, max_value=12), seed=st.integers(0, 10_000),
       skew=st.floats(min_value=0.0, max_value=6.0))
def test_top1gating_invariants(tokens, experts, seed, skew):
    # tokens >= min_capacity (4): below that DeepSpeed's topk(k = capacity) itself raises, as the restatement does
    g = torch.Generator().manual_seed(seed)
    logits = torch.randn(tokens, experts, generator=g)
    logits[:, 0] += skew  # push load onto expert 0 so that capacity overflow / dropping is exercised
    uniform = torch.rand(tokens, experts, generator=g)
    C = O.moe_capacity(tokens, experts)
    l_aux, combine, dispatch, counts, (idx, slot, prob) = O.top1gating(logits, C, uniform)
    assert C == max(-(-tokens // experts), 4)
    assert int(counts.sum()) == tokens and torch.equal(counts, torch.bincount(idx, minlength=experts))
    kept = slot >= 0
    for e in range(experts):
        mine = (idx == e)
        n_kept = int((mine & kept).sum())
        assert n_kept == min(int(counts[e]), C)                       # an expert drops tokens only when over capacity
        s = slot[mine & kept]
        assert sorted(s.tolist()) == list(range(n_kept))                # slots are 0..n-1, each used once
        assert s.tolist() == sorted(s.tolist())                         # and follow TOKEN ORDER (cumsum of the kept mask)
        if int(counts[e]) > C:                                          # Random-Token-Selection: the C largest uniform draws stay
            u = uniform[mine, e]
            thr = torch.topk(u, C).values.min()
            assert torch.equal(kept[mine], u >= thr)
    # dense tensors the reference builds == the sparse view the CUDA path produces
    assert dispatch.sum().item() == int(kept.sum()) and combine.shape == (tokens, experts, C)
    t = torch.arange(tokens)[kept]
    assert torch.allclose(combine[t, idx[kept], slot[kept]], prob[kept])
    assert torch.allclose(combine.sum(dim=(1, 2)), torch.where(kept, prob, torch.zeros_like(prob)))
    me, ce = torch.softmax(logits, 1).mean(0), torch.nn.functional.one_hot(idx, experts).float().mean(0)
    assert torch.allclose(l_aux, (me * ce).sum() * experts)


def test_weave_schedule_integer_exact():
    """cn_block_idx = int(i / (n_base / n_ctrl)) (src/UniGenTransformer.py:1126-1127,1159-1160)."""
    assert O.weave_schedule(19, 9) == [0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8]
    assert O.weave_schedule(38, 19) == [i // 2 for i in range(38)]
    assert O.weave_schedule(24, 24) == list(range(24))
