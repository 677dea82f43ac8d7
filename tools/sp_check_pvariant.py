"""Sequence-parallel P-variant (segment-sharded Ulysses over NVLink peer memory) vs the single-GPU P-variant forward on
identical weights / inputs — BASELINE cfg4 in its "~17 k tokens" reading (T + N + 3 Nc = 16 896 tokens at 1024^2).
torchrun --nproc-per-node P tools/sp_check_pvariant.py [--workload tiny|cfg4] [--conds 3] [--steps K] [--graph]
-> one JSON line from rank 0."""
import argparse
import json
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="tiny")
    ap.add_argument("--conds", type=int, default=3)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--graph", action="store_true")
    ap.add_argument("--skip-single", action="store_true", help="time only the sequence-parallel run (no single-GPU comparison)")
    args = ap.parse_args()
    from unigen_b200 import ops
    from unigen_b200.model import FluxArch
    from unigen_b200.parallel import SequenceParallelUniCombineFlux
    from unigen_b200.pvariant import DOUBLE_LORA, SINGLE_LORA, UniCombineFlux
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    tiny = args.workload == "tiny"
    arch = FluxArch.tiny() if tiny else FluxArch()
    side, T = (256, 64) if tiny else (1024, 512)
    grid = side // 16
    N, D = grid * grid, arch.num_attention_heads * arch.attention_head_dim
    types_ = ["depth", "canny", "openpose"][:args.conds]
    adapters = ["denoise"] + types_
    sp = SequenceParallelUniCombineFlux(arch, device=dev, lora_rank=4, max_conditions=args.conds)
    g = torch.Generator(device=dev).manual_seed(0)
    with torch.no_grad():
        for k, v in sp._ws.views.items():
            if "norm_q" in k or "norm_k" in k or "norm_added" in k:
                v.fill_(1.0)
            else:
                fan = sp._ws.views[k[:-4] + "weight"].shape[-1] if k.endswith(".bias") else v.shape[-1]
                v.copy_((torch.rand(v.shape, device=dev, generator=g) * 2 - 1) / fan ** 0.5)
    sd = {}
    names = ["x_embedder"] + [f"transformer_blocks.{i}.{n}" for i in range(arch.num_layers) for n in DOUBLE_LORA] + \
            [f"single_transformer_blocks.{i}.{n}" for i in range(arch.num_single_layers) for n in SINGLE_LORA]
    for name in names:
        out_f, in_f = sp._ws.views[name + ".weight"].shape
        for a in adapters:
            sd[f"{name}.lora_A.{a}.weight"] = torch.randn(4, in_f, device=dev, generator=g) / in_f ** 0.5
            sd[f"{name}.lora_B.{a}.weight"] = torch.randn(out_f, 4, device=dev, generator=g) * 0.25
    sp.load_state_dict(sd, adapters=adapters, condition_types=types_)
    del sd
    ids = torch.zeros(grid, grid, 3, device=dev)
    ids[..., 1] += torch.arange(grid, device=dev)[:, None]
    ids[..., 2] += torch.arange(grid, device=dev)[None, :]
    ids = ids.reshape(N, 3)
    gi = torch.Generator(device=dev).manual_seed(7)
    rnd = lambda *s: torch.randn(*s, device=dev, generator=gi).to(torch.bfloat16)  # noqa: E731
    cond_ids = []
    for j in range(args.conds):  # distinct positions per condition so a mis-routed RoPE row cannot hide
        c = ids.clone()
        c[:, 2] += (j + 1) * grid
        cond_ids.append(c)
    argsf = (rnd(1, N, 64), [rnd(1, N, 64) for _ in types_], cond_ids, types_, rnd(1, T, 4096),
             torch.randn(1, 768, device=dev, generator=gi), torch.tensor([0.5], device=dev), ids, torch.zeros(T, 3, device=dev))
    out_sp = sp(*argsf).float().clone()
    rel, ms_ref = None, None
    if not args.skip_single:
        ref = UniCombineFlux(arch, device=dev, lora_rank=4, max_conditions=args.conds)
        ref._ws = sp._ws
        for name in ("x_embedder_w", "context_embedder_w", "time_text", "double", "single", "norm_out_w", "proj_out_w", "lora", "R",
                     "condition_types"):
            setattr(ref, name, getattr(sp, name))
        out_ref = ref(*argsf).float()
        rel = ((out_sp - out_ref).norm() / out_ref.norm()).item()

    def timed(fn):
        for _ in range(2):
            fn()
        dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            fn()
        e1.record()
        dist.barrier(); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / args.steps], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    graph_equal = None
    if args.graph:
        sp.use_cuda_graph = True
        for _ in range(2):
            out_g = sp(*argsf).float()
        graph_equal = bool(torch.equal(out_g, out_sp))
    ops.reset_launch_count()
    ms_sp = timed(lambda: sp(*argsf))
    launches = ops.launch_count() // (args.steps + 2)
    if not args.skip_single:
        ms_ref = timed(lambda: ref(*argsf))
    errs = torch.tensor([sp._pool.error()], device=dev)
    dist.all_reduce(errs, op=dist.ReduceOp.MAX)
    S = T + N + args.conds * N
    pairs = (T + N) * S + args.conds * N * (T + N + N)
    flop = (2 * 12 * S * D * D + 4 * pairs * D) * (arch.num_layers + arch.num_single_layers)
    if rank == 0:
        ok = errs.item() == 0 and graph_equal is not False and (rel is None or rel < 5e-3) and bool(torch.isfinite(out_sp).all())
        print(json.dumps({"check": "segment_sharded_ulysses_pvariant_vs_single_gpu", "workload": args.workload, "tokens": S, "world": world,
                          "rel_l2": rel, "cuda_graph": bool(args.graph), "graph_equals_eager": graph_equal,
                          "peer_barrier_timeouts": int(errs.item()), "ok": bool(ok), "ms_per_step_sp": ms_sp,
                          "ms_per_step_single_gpu": ms_ref, "latency_speedup": (ms_ref / ms_sp) if ms_ref else None,
                          "tflop_per_step": flop / 1e12, "model_tflops_aggregate": flop / ms_sp / 1e9,
                          "gpu_launches_per_step_per_rank": launches}), flush=True)
    sp._pool.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
