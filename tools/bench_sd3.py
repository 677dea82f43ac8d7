"""cfg5: SD3.5-medium-architecture UniGenSD3 (24 MMDiT + 24 control blocks, hidden 1536, 13 dual-attention layers,
transformer-block CoMoE experts) at 1024x1024 + 1 condition: 4096 image + 4096 condition + 333 text tokens.
BASELINE cfg5 runs batch 16 across 8 B200 with CFG, i.e. per-GPU micro-batch 2 and a CFG-doubled forward batch of 4.

python tools/bench_sd3.py [--batch 4] [--steps 5] [--side 1024] [--no-graph]   -> one JSON line
Under torchrun: every rank runs its own micro-batch (data parallel, no data-path collective), value = max over ranks."""
import argparse
import json
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))


def sd3_step_flops(D, N, T, L, n_dual_base, n_dual_ctrl, E, C, in_k=64, joint=4096, pooled=2048, executed=True):
    """FLOPs (2 x MAC) of one UniGenSD3 forward per sample. Joint block: sample rows 12 D^2 (+4 D^2 with attn2), context
    rows 12 D^2 (3 D^2 when its post-attention half is skipped / context_pre_only); attention 4 S^2 D (+ 4 n^2 D for attn2).
    executed=True leaves out the text-stream post-attention half of the control blocks (their text output is discarded)."""
    def joint_blk(n_smp, n_ctx, dual, ctx_post):
        g = n_smp * (12 + (4 if dual else 0)) * D * D + n_ctx * (12 if ctx_post else 3) * D * D
        a = 2 * (n_smp + n_ctx) ** 2 * D + (2 * n_smp ** 2 * D if dual else 0)
        return 2 * g, 2 * a

    g = a = 0
    for i in range(L):
        x, y = joint_blk(N, T, i < n_dual_base, i != L - 1); g += x; a += y
        x, y = joint_blk(N, T, i < n_dual_ctrl, not executed); g += x + 2 * N * D * D; a += y
    g += 2 * (2 * N * in_k * D + T * joint * D + T * D * D + N * D * E + N * D * in_k)
    x, y = joint_blk(N, N, False, True); g += x; a += y                 # shared_expert[0]
    x, y = joint_blk(2 * N, T, True, False); g += x; a += y             # shared_expert[1] (context_pre_only)
    g += 2 * 2 * E * C * 12 * D * D; a += 2 * 2 * E * 2 * C * C * D     # 2 branches x E single blocks over C slots
    return g, a


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--side", type=int, default=1024)
    ap.add_argument("--text", type=int, default=333)
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--attn-variant", type=int, default=0)
    ap.add_argument("--gemm-variant", type=int, default=0)
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from unigen_b200 import ops
    from unigen_b200.sd3 import SD3Arch, UniGenSD3, shipped_control_params
    arch = SD3Arch()
    model = UniGenSD3(arch, device=f"cuda:{local}")
    model.init_condition_block(condition_nums=1, control_params=shipped_control_params())
    model.init_random_(seed=0)
    model.use_cuda_graph = not args.no_graph
    model.attn_variant, model.gemm_variant = args.attn_variant, args.gemm_variant
    B, lat = args.batch, args.side // 8
    N, T, D = (lat // 2) ** 2, args.text, model.inner_dim
    g = torch.Generator().manual_seed(1234 + rank)
    host = dict(hidden_states=torch.randn(B, 16, lat, lat, generator=g).bfloat16().pin_memory(),
                condition_hidden_states=torch.randn(B, 16, lat, lat, generator=g).bfloat16().pin_memory(),
                encoder_hidden_states=torch.randn(B, T, 4096, generator=g).bfloat16().pin_memory(),
                pooled_projections=torch.randn(B, 2048, generator=g).pin_memory(),
                condition_pooled_projections=torch.randn(B, 2048, generator=g).pin_memory(),
                timestep=torch.full((B,), 500.0).pin_memory(),
                rts_uniform=torch.rand(B * N, model.expert_nums, generator=g).pin_memory())
    dev = {k: v.cuda(non_blocking=True) for k, v in host.items()}

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(args.warmup):
        out = model(**dev)[0]
    sync()
    ops.reset_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        out = model(**dev)[0]
    e1.record()
    sync()
    ms = e0.elapsed_time(e1) / args.steps
    launches = ops.launch_count()
    # end to end: pinned host inputs -> device, forward, velocity back to the host, every step
    e0.record()
    for _ in range(args.steps):
        d = {k: v.cuda(non_blocking=True) for k, v in host.items()}
        vel = model(**d)[0].to("cpu", non_blocking=True)
    e1.record()
    sync()
    ms_e2e = e0.elapsed_time(e1) / args.steps
    if world > 1:
        t = torch.tensor([ms, ms_e2e], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = t.tolist()
    C = ops.moe_capacity(B * N, model.expert_nums)
    gf, af = sd3_step_flops(D, N, T, arch.num_layers, 13, 13, model.expert_nums, C / B)
    gr, ar = sd3_step_flops(D, N, T, arch.num_layers, 13, 13, model.expert_nums, C / B, executed=False)
    if rank == 0:
        print(json.dumps({
            "workload": f"cfg5 architecture: SD3.5-medium UniGenSD3 {args.side}x{args.side} + 1 condition ({N} image + {N} condition + {T} text tokens), "
                        f"forward batch {B} per GPU ({'CFG-doubled micro-batch ' + str(B // 2) if B % 2 == 0 else 'no CFG'})",
            "n_gpus": world, "ms_per_step": ms, "sample_steps_per_s": world * B * 1e3 / ms,
            "e2e": {"ms_per_step": ms_e2e, "sample_steps_per_s": world * B * 1e3 / ms_e2e,
                    "h2d_bytes_per_step": sum(v.numel() * v.element_size() for v in host.values()), "d2h_bytes_per_step": vel.numel() * 2},
            "tflop_per_step_per_sample_executed": (gf + af) / 1e12, "tflop_per_step_per_sample_reference_algorithmic": (gr + ar) / 1e12,
            "model_tflops_per_gpu": B * (gf + af) / ms / 1e9, "gpu_launches_per_step": launches // args.steps,
            "cuda_graph": model.use_cuda_graph, "attn_variant": args.attn_variant, "gemm_variant": args.gemm_variant, "finite": bool(torch.isfinite(out.float()).all())}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
