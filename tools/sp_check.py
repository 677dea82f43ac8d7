"""Sequence-parallel (Ulysses) forward vs the single-GPU forward on identical weights / inputs.
torchrun --nproc-per-node P tools/sp_check.py [--workload tiny|cfg3] [--steps K]  -> one JSON line from rank 0."""
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
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"])
    ap.add_argument("--graph", action="store_true", help="replay the sequence-parallel step as one CUDA graph per rank (peer exchange)")
    args = ap.parse_args()
    from unigen_b200.model import FluxArch, UniGenFlux, canonical_control_params
    from unigen_b200.parallel import SequenceParallelUniGenFlux
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    tiny = args.workload == "tiny"
    arch = FluxArch.tiny() if tiny else FluxArch()
    side, T = (256, 512) if tiny else (1024, 512)
    N = (side // 16) ** 2
    g = torch.Generator().manual_seed(7)
    grid = side // 16
    ids = torch.zeros(grid, grid, 3)
    ids[..., 1] += torch.arange(grid)[:, None]
    ids[..., 2] += torch.arange(grid)[None, :]
    ids = ids.reshape(N, 3)
    E = 6
    inp = dict(hidden_states=torch.randn(1, N, 64, generator=g), condition_hidden_states=torch.randn(1, N, 64, generator=g),
               encoder_hidden_states=torch.randn(1, T, 4096, generator=g), pooled_projections=torch.randn(1, 768, generator=g),
               condition_pooled_projections=torch.randn(1, 768, generator=g), timestep=torch.tensor([0.5]), img_ids=ids,
               txt_ids=torch.zeros(T, 3), condition_ids=ids.clone(), rts_uniform=torch.rand(N, E, generator=g))
    inp = {k: v.to(dev) for k, v in inp.items()}
    sp = SequenceParallelUniGenFlux(arch, device=dev, exchange=args.exchange)
    sp.init_condition_block(condition_nums=1, control_params=canonical_control_params())
    sp.init_random_(seed=0)
    out_sp = sp(**inp)[0].float().clone()
    # single-GPU reference on the SAME weights (views of the same storage)
    ref = UniGenFlux(arch, device=dev)
    ref.init_condition_block(condition_nums=1, control_params=canonical_control_params())
    ref._ws = sp._ws
    for name in vars(sp):
        if name.endswith("_w") or name in ("time_text", "control_time_text", "control_condition", "double", "single", "ctrl_double",
                                           "ctrl_single", "add_double", "add_single", "gate_wg", "exp_w", "exp_b", "exp_mod_w",
                                           "exp_mod_b", "shared"):
            setattr(ref, name, getattr(sp, name))
    out_ref = ref(**inp)[0].float()
    rel = ((out_sp - out_ref).norm() / out_ref.norm()).item()
    same_route = bool(torch.equal(sp._last_route["slot"], ref._last_route["slot"]))
    # timing
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
    sp.use_cuda_graph = ref.use_cuda_graph = bool(args.graph)
    if args.graph:  # the graph replay must reproduce the eager sequence-parallel result bit for bit
        for _ in range(2):
            out_g = sp(**inp)[0].float()
        graph_equal = bool(torch.equal(out_g, out_sp))
    else:
        graph_equal = None
    ms_sp = timed(lambda: sp(**inp))
    ms_ref = timed(lambda: ref(**inp))
    peer_err = sp._pool.error() if sp._pool is not None else 0
    errs = torch.tensor([peer_err], device=dev)
    dist.all_reduce(errs, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({"check": "ulysses_sp_vs_single_gpu", "workload": args.workload, "world": world, "rel_l2": rel,
                          "exchange": args.exchange,
                          "cuda_graph": bool(args.graph), "graph_equals_eager": graph_equal, "peer_barrier_timeouts": int(errs.item()),
                          "routing_identical": same_route,
                          "ok": bool(rel < 5e-3 and same_route and errs.item() == 0 and graph_equal is not False), "ms_per_step_sp": ms_sp,
                          "ms_per_step_single_gpu": ms_ref, "latency_speedup": ms_ref / ms_sp}), flush=True)
    if sp._pool is not None:
        sp._pool.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
