"""LN-modulate kernels A/B (UG_LN_KERNEL=stream|rows|warp, one subprocess each): timing on the cfg3 shapes and a checksum of
the output (stream and rows must be bit-identical)."""
import hashlib
import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def child():
    import torch
    from unigen_b200 import ops
    dev = torch.device("cuda")
    g = torch.Generator(device=dev).manual_seed(0)
    rec = {"kernel": os.environ.get("UG_LN_KERNEL", "default")}
    for B, R, D in ((1, 4608, 3072), (1, 4096, 3072), (1, 512, 3072), (8, 4608, 3072), (2, 4429, 1536), (1, 577, 3072)):
        x = (torch.randn(B, R, D, device=dev, generator=g) * 2).to(torch.bfloat16)
        shift, scale = torch.randn(B, D, device=dev, generator=g), torch.randn(B, D, device=dev, generator=g)
        out = torch.empty_like(x)
        big = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        for _ in range(3):
            ops.ln_modulate(x, out, shift, scale)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            ops.ln_modulate(x, out, shift, scale)
        e1.record()
        torch.cuda.synchronize()
        warm = e0.elapsed_time(e1) / 20 * 1e3
        cold = []
        for _ in range(5):
            big.fill_(1)  # flush L2
            e0.record(); ops.ln_modulate(x, out, shift, scale); e1.record()
            torch.cuda.synchronize()
            cold.append(e0.elapsed_time(e1) * 1e3)
        rec[f"{B}x{R}x{D}"] = {"warm_us": round(warm, 1), "cold_us": round(sorted(cold)[2], 1),
                               "warm_TBps": round(4.0 * B * R * D / warm / 1e6, 2),
                               "sha": hashlib.sha1(out.view(torch.int16).cpu().numpy().tobytes()).hexdigest()[:12]}
    print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--child":
        child()
    else:
        for k in ("stream", "rows", "warp"):
            r = subprocess.run([sys.executable, __file__, "--child"], env=dict(os.environ, UG_LN_KERNEL=k), capture_output=True, text=True)
            print(r.stdout.strip() or r.stderr[-2000:])
