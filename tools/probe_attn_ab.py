"""Same-box A/B of attention variants: bit-identity against variant 5 and interleaved timing (CUDA events, 20 launches per arm, 3 rounds).
    python tools/probe_attn_ab.py [variants ...]     default: 5 8"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from unigen_b200 import ops  # noqa: E402

variants = [int(v) for v in sys.argv[1:]] or [5, 8]
dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(0)
cases = [("cfg3 joint S=4608 H=24 dh=128", 1, 4608, 24, 128, None, None),
         ("pre-stage S=8704 H=24 dh=128", 1, 8704, 24, 128, None, None),
         ("P-variant S=16896 3 cond segments", 1, 16896, 24, 128, [0, 512, 4608, 8704, 12800, 16896], [0b11111, 0b11111, 0b00111, 0b01011, 0b10011]),
         ("SD3.5 S=4429 B=4 H=24 dh=64", 4, 4429, 24, 64, None, None),
         ("ragged S=1000 H=6 dh=128 2 segments", 2, 1000, 6, 128, [0, 300, 1000], [0b11, 0b10])]
for name, B, S, H, dh, segs, vis in cases:
    D = H * dh
    qkv = (torch.randn(B, S, 3 * D, device=dev, generator=g) * 1.5).to(torch.bfloat16)
    q, k, v = qkv[:, :, :D], qkv[:, :, D:2 * D], qkv[:, :, 2 * D:]
    outs, times = {}, {x: [] for x in variants}
    for x in variants:
        o = torch.zeros(B, S, D, device=dev, dtype=torch.bfloat16)
        ops.attention(q, k, v, o, H, dh, seg_bounds=segs, seg_visible=vis, variant=x)
        outs[x] = o
    torch.cuda.synchronize()
    for _ in range(3):
        for x in variants:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                ops.attention(q, k, v, outs[x], H, dh, seg_bounds=segs, seg_visible=vis, variant=x)
            e1.record()
            torch.cuda.synchronize()
            times[x].append(e0.elapsed_time(e1) / 20 * 1e3)
    fl = 4.0 * B * H * S * S * dh  # dense count (masked tiles are skipped: the P-variant line over-counts)
    base = outs[variants[0]]
    for x in variants:
        us = min(times[x])
        print(f"{name}: variant {x}: {us:9.1f} us  {fl / us / 1e6:7.1f} TFLOP/s (dense count)  bit-identical to variant {variants[0]}: "
              f"{torch.equal(outs[x], base)}  nan: {bool(torch.isnan(outs[x].float()).any())}", flush=True)
