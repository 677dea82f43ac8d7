"""Times the staged-epilogue GEMM variants (4 = 2-CTA 256x256, 5 = 1-CTA 128x256, 6 = 1-CTA 128x128) and the auto choice on the
per-rank shapes of sequence parallelism (M = rows / P) and the full shapes; prints TFLOP/s per variant (isolated, L2-warm)."""
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from unigen_b200 import ops  # noqa: E402


def main():
    dev = torch.device("cuda")
    shapes = []
    for M in (576, 1152, 2112, 2304, 4608):
        for N, K in ((9216, 3072), (12288, 3072), (3072, 15360), (3072, 3072), (3072, 12288)):
            shapes.append((M, N, K))
    for M, N, K in shapes:
        a = torch.randn(1, M, K, device=dev).to(torch.bfloat16)
        w = torch.randn(N, K, device=dev).to(torch.bfloat16) * 0.02
        bias = torch.randn(N, device=dev).to(torch.bfloat16)
        out = torch.empty(1, M, N, device=dev, dtype=torch.bfloat16)
        rec = {"M": M, "N": N, "K": K}
        for v in (0, 4, 5, 6):
            try:
                for _ in range(3):
                    ops.gemm(a, w, out=out, bias=bias, variant=v)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                iters = 20
                e0.record()
                for _ in range(iters):
                    ops.gemm(a, w, out=out, bias=bias, variant=v)
                e1.record()
                torch.cuda.synchronize()
                us = e0.elapsed_time(e1) / iters * 1e3
                rec[f"v{v}_us"] = round(us, 1)
                rec[f"v{v}_tf"] = round(2.0 * M * N * K / us / 1e6, 0)
            except Exception as e:  # noqa: BLE001
                rec[f"v{v}"] = str(e)[:80]
        best = min((4, 5, 6), key=lambda v: rec.get(f"v{v}_us", 1e9))
        rec["best"] = best
        rec["auto_vs_best"] = round(rec["v0_us"] / rec[f"v{best}_us"], 3)
        print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()
