"""Per-op device-time breakdown of ONE eager sequence-parallel step on every rank (CUDA events around each C-ABI call).
Event pairs serialise nothing, but an eager step is launch-bound at small per-rank work: read the SHARES, not the sum.
torchrun --nproc-per-node P tools/sp_profile.py [--model s|p] [--workload tiny|full]  -> one JSON line per reporting rank."""
import argparse
import json
import os
import sys
from collections import defaultdict
from pathlib import Path

import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

OPS = ["gemm", "attention", "attention_peer", "qkv_scatter", "peer_barrier", "peer_bcast_rows", "ln_modulate", "qk_rmsnorm_rope",
       "gemv", "lora_down", "lora_down_wide", "ln_modulate_segs", "add", "copy", "moe_route", "moe_gather_modulate", "moe_combine", "rope_table", "to_bf16"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="s", choices=["s", "p"])
    ap.add_argument("--workload", default="full")
    args = ap.parse_args()
    from unigen_b200 import ops
    from unigen_b200.model import FluxArch, canonical_control_params
    from unigen_b200.parallel import SequenceParallelUniCombineFlux, SequenceParallelUniGenFlux
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    tiny = args.workload == "tiny"
    arch = FluxArch.tiny() if tiny else FluxArch()
    side, T = (256, 64 if args.model == "p" else 512) if tiny else (1024, 512)
    grid = side // 16
    N = grid * grid
    g = torch.Generator(device=dev).manual_seed(0)
    ids = torch.zeros(grid, grid, 3, device=dev)
    ids[..., 1] += torch.arange(grid, device=dev)[:, None]
    ids[..., 2] += torch.arange(grid, device=dev)[None, :]
    ids = ids.reshape(N, 3)
    rnd = lambda *s: torch.randn(*s, device=dev, generator=g).to(torch.bfloat16)  # noqa: E731
    if args.model == "s":
        m = SequenceParallelUniGenFlux(arch, device=dev, exchange="peer")
        m.init_condition_block(condition_nums=1, control_params=canonical_control_params())
        m.init_random_(seed=0)
        inp = dict(hidden_states=rnd(1, N, 64), condition_hidden_states=rnd(1, N, 64), encoder_hidden_states=rnd(1, T, 4096),
                   pooled_projections=torch.randn(1, 768, device=dev, generator=g),
                   condition_pooled_projections=torch.randn(1, 768, device=dev, generator=g), timestep=torch.tensor([0.5], device=dev),
                   img_ids=ids, txt_ids=torch.zeros(T, 3, device=dev), condition_ids=ids.clone(),
                   rts_uniform=torch.rand(N, 6, device=dev, generator=g))
        step = lambda: m(**inp)  # noqa: E731
    else:
        types_ = ["depth", "canny", "openpose"]
        m = SequenceParallelUniCombineFlux(arch, device=dev, lora_rank=4, max_conditions=3)
        with torch.no_grad():
            for k, v in m._ws.views.items():
                v.copy_((torch.rand(v.shape, device=dev, generator=g) * 2 - 1) * 0.02) if "norm_" not in k else v.fill_(1.0)
        m.load_state_dict({}, adapters=["denoise"] + types_, condition_types=types_)
        argsf = (rnd(1, N, 64), [rnd(1, N, 64) for _ in types_], [ids.clone() for _ in types_], types_, rnd(1, T, 4096),
                 torch.randn(1, 768, device=dev, generator=g), torch.tensor([0.5], device=dev), ids, torch.zeros(T, 3, device=dev))
        step = lambda: m(*argsf)  # noqa: E731
    for _ in range(2):
        step()
    torch.cuda.synchronize()
    dist.barrier()
    records = []
    orig = {}

    def wrap(name):
        fn = getattr(ops, name)
        orig[name] = fn

        def timed(*a, **k):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = fn(*a, **k)
            e1.record()
            records.append((name, e0, e1))
            return out
        setattr(ops, name, timed)

    for name in OPS:
        if hasattr(ops, name):
            wrap(name)
    m.overlap_mod_gemv = False  # one stream: events of the side stream would not be comparable
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    step()
    t1.record()
    torch.cuda.synchronize()
    for name, fn in orig.items():
        setattr(ops, name, fn)
    tot, cnt = defaultdict(float), defaultdict(int)
    for name, e0, e1 in records:
        tot[name] += e0.elapsed_time(e1)
        cnt[name] += 1
    rec = {"rank": rank, "world": world, "model": args.model, "eager_step_ms": t0.elapsed_time(t1),
           "ops_ms": {k: round(v, 3) for k, v in sorted(tot.items(), key=lambda kv: -kv[1])}, "ops_n": dict(cnt)}
    for r in (0, 1, world - 1):
        if rank == r:
            print(json.dumps(rec), flush=True)
        dist.barrier()
    m._pool.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
