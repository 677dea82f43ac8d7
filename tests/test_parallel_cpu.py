"""World-size-2 gloo tests (CPU) of the multi-GPU host logic: data-parallel sharding and the Ulysses all-to-all layout
around attention. The exchange code is the product code of unigen_b200/parallel.py; only the strided-copy primitive and
the attention stand-in are torch here (the CUDA kernels need a GPU)."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def test_dp_shard_covers_all_samples_once():
    from unigen_b200.parallel import dp_shard
    for n, world in [(8, 1), (8, 2), (8, 8), (7, 4), (3, 8)]:
        got = [i for r in range(world) for i in dp_shard(n, world, r)]
        assert got == list(range(n))


def test_sp_row_split():
    from unigen_b200.parallel import sp_row_split
    T, N = 512, 4096
    for world in (1, 2, 4, 8):
        rows_seen, txt, img0 = 0, 0, []
        for r in range(world):
            row0, rows, t_loc, i0 = sp_row_split(T, N, world, r)
            assert row0 == rows_seen and rows == (T + N) // world
            rows_seen += rows
            txt += t_loc
            img0.append(i0)
        assert rows_seen == T + N and txt == T and img0[0] == 0
    with pytest.raises(ValueError):
        sp_row_split(511, 4096, 2, 0)


def _worker(rank, world, port, S, H, dh, ret):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import unigen_oracle as O
        from unigen_b200.parallel import UlyssesExchange
        torch.manual_seed(0)
        D = H * dh
        qkv = torch.randn(S, 3 * D)                      # identical on every rank (same seed)
        s_loc = S // world
        big = torch.zeros(s_loc, 5 * D)                  # output lives inside a wider buffer (row stride 5D, like CAT)
        x = UlyssesExchange(None, world, s_loc, D, "cpu", torch.float32, copy=lambda s, d: d.copy_(s))
        q, k, v = x.seq_to_heads(qkv[rank * s_loc:(rank + 1) * s_loc])
        hp = H // world
        o = O.sdpa(*(t.reshape(1, S, hp, dh).transpose(1, 2) for t in (q, k, v)))   # this rank's heads, all tokens
        x.o_full.copy_(o.transpose(1, 2).reshape(S, hp * dh))
        out = x.heads_to_seq(big[:, :D])
        full = O.sdpa(*(qkv[:, i * D:(i + 1) * D].reshape(1, S, H, dh).transpose(1, 2) for i in range(3)))
        want = full.transpose(1, 2).reshape(S, D)[rank * s_loc:(rank + 1) * s_loc]
        ret[rank] = float((out - want).abs().max())
        assert big[:, D:].abs().max() == 0
    finally:
        dist.destroy_process_group()


def test_ulysses_exchange_equals_unsharded_attention_world2():
    world, S, H, dh = 2, 48, 6, 8
    with mp.Manager() as m:
        ret = m.dict()
        mp.spawn(_worker, args=(world, 29533, S, H, dh, ret), nprocs=world, join=True)
        assert len(ret) == world and max(ret.values()) < 1e-5, dict(ret)


def test_pool_layout_is_aligned_ordered_and_deterministic():
    """Peer pools must have IDENTICAL offsets on every rank: the layout is a pure function of the size list."""
    from unigen_b200._lib import UG_PEER_HEADER_BYTES
    from unigen_b200.parallel import pool_layout
    sizes = [("RECV", 3 * 8704 * 384 * 2), ("AO", 8704 * 3072 * 2), ("CAT", 4608 * 15360 * 2), ("X", 4608 * 3072 * 2), ("OUTF", 4608 * 64 * 2 + 2)]
    off, total = pool_layout(sizes)
    assert pool_layout(sizes) == (off, total)
    prev_end = UG_PEER_HEADER_BYTES
    for name, nbytes in sizes:
        assert off[name] % 256 == 0 and off[name] >= prev_end
        prev_end = off[name] + nbytes
    assert total >= prev_end and total % 256 == 0
