"""cfg4, P-variant reading: Flux-arch LoRA-switched joint blocks with ALL condition tokens in every block
(T + N + 3 Nc = 16 896 tokens at 1024^2, SURVEY.md §8d cfg4: 347.6 TFLOP / step) on one B200.
python tools/bench_pvariant.py [--conds 3] [--steps 3] [--side 1024]   -> one JSON line"""
import argparse
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--conds", type=int, default=3)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--side", type=int, default=1024)
    ap.add_argument("--strict", action="store_true")
    ap.add_argument("--lora-mode", default="mma", choices=["mma", "epilogue"])
    args = ap.parse_args()
    from unigen_b200 import ops
    from unigen_b200.model import FluxArch
    from unigen_b200.pvariant import DOUBLE_LORA, SINGLE_LORA, UniCombineFlux
    arch = FluxArch()
    D, T, N = 3072, 512, (args.side // 16) ** 2
    types_ = ["depth", "canny", "openpose"][:args.conds]
    adapters = ["denoise"] + types_
    model = UniCombineFlux(arch, device="cuda", lora_rank=4, max_conditions=args.conds, strict_mask=args.strict)
    g = torch.Generator(device="cuda").manual_seed(0)
    with torch.no_grad():
        for k, v in model._ws.views.items():
            if "norm_q" in k or "norm_k" in k or "norm_added" in k:
                v.fill_(1.0)
            else:
                fan = model._ws.views[k[:-4] + "weight"].shape[-1] if k.endswith(".bias") else v.shape[-1]
                v.copy_((torch.rand(v.shape, device="cuda", generator=g) * 2 - 1) / fan ** 0.5)
    # random LoRA pairs for every switched linear (PEFT names), then the normal loader builds the stacked groups
    sd = {}
    names = ["x_embedder"] + [f"transformer_blocks.{i}.{n}" for i in range(arch.num_layers) for n in DOUBLE_LORA] + \
            [f"single_transformer_blocks.{i}.{n}" for i in range(arch.num_single_layers) for n in SINGLE_LORA]
    for name in names:
        out_f, in_f = model._ws.views[name + ".weight"].shape
        for a in adapters:
            sd[f"{name}.lora_A.{a}.weight"] = torch.randn(4, in_f, device="cuda", generator=g) / in_f ** 0.5
            sd[f"{name}.lora_B.{a}.weight"] = torch.randn(out_f, 4, device="cuda", generator=g) * 0.25
    model.load_state_dict(sd, adapters=adapters, condition_types=types_)
    model.lora_mode = args.lora_mode
    grid = args.side // 16
    ids = torch.zeros(grid, grid, 3, device="cuda")
    ids[..., 1] += torch.arange(grid, device="cuda")[:, None]
    ids[..., 2] += torch.arange(grid, device="cuda")[None, :]
    ids = ids.reshape(N, 3)
    rnd = lambda *s: torch.randn(*s, device="cuda", generator=g).to(torch.bfloat16)  # noqa: E731
    argsf = (rnd(1, N, 64), [rnd(1, N, 64) for _ in types_], [ids.clone() for _ in types_], types_, rnd(1, T, 4096),
             torch.randn(1, 768, device="cuda", generator=g), torch.tensor([0.5], device="cuda"), ids, torch.zeros(T, 3, device="cuda"))
    for _ in range(2):
        model(*argsf)
    torch.cuda.synchronize()
    ops.reset_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        out = model(*argsf)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    S = T + N + args.conds * N
    Sc = N  # each condition stream
    # GEMM: every token passes 12 D^2 MACs per block; attention with the reference visibility rule:
    # main queries (T+N) see all S keys, each condition's Nc queries see T+N+Nc keys
    pairs = (T + N) * S + args.conds * Sc * (T + N + Sc) if not args.strict else (T + N) * S + args.conds * Sc * Sc
    gemm = 2 * 12 * S * D * D * (arch.num_layers + arch.num_single_layers)
    attn = 4 * pairs * D * (arch.num_layers + arch.num_single_layers)
    print(json.dumps({"workload": f"cfg4 P-variant: {S} tokens ({args.conds} conditions), LoRA rank 4 switched per segment, "
                                  f"{'strict' if args.strict else 'reference'} visibility mask",
                      "ms_per_step": ms, "steps_per_s": 1e3 / ms, "tflop_per_step": (gemm + attn) / 1e12,
                      "model_tflops": (gemm + attn) / ms / 1e9, "gpu_launches_per_step": ops.launch_count() // args.steps, "lora_mode": args.lora_mode,
                      "finite": bool(torch.isfinite(out.float()).all())}))


if __name__ == "__main__":
    main()
