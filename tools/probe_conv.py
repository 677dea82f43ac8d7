"""Same-box timing of the implicit 3x3 convolution's tile variants at the VAE's shapes (CUDA events, 20 back-to-back launches)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from unigen_b200 import ops  # noqa: E402

dev = torch.device("cuda")
bf = torch.bfloat16
g = torch.Generator(device=dev).manual_seed(0)
for (H, W, Ci, Co) in ((1024, 1024, 128, 128), (512, 512, 256, 256), (512, 512, 256, 128), (256, 256, 512, 512), (1024, 1024, 128, 64)):
    x = torch.randn(1, H, W, Ci, device=dev, generator=g).to(bf)
    w = (torch.randn(Co, 9 * Ci, device=dev, generator=g) / (3 * Ci ** 0.5)).to(bf)
    b = torch.zeros(Co, device=dev, dtype=bf)
    y = torch.empty(1, H, W, Co, device=dev, dtype=bf)
    fl = 2.0 * H * W * 9 * Ci * Co
    ref = None
    for v in (0, 4, 5, 6, 7):
        if v == 7 and Co > 128:
            continue
        try:
            for _ in range(3):
                ops.conv3x3(x, w, bias=b, residual=x if Ci == Co else None, out=y, variant=v)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                ops.conv3x3(x, w, bias=b, residual=x if Ci == Co else None, out=y, variant=v)
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) / 20 * 1e3
            same = "" if ref is None else f" equal_to_first={torch.equal(ref, y)}"
            if ref is None:
                ref = y.clone()
            print(f"{H}x{W} {Ci}->{Co} variant {v}: {us:8.1f} us  {fl / us / 1e6:7.1f} TFLOP/s{same}", flush=True)
        except Exception as e:  # noqa: BLE001
            print(f"{H}x{W} {Ci}->{Co} variant {v}: {e}", flush=True)
