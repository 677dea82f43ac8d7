"""Run one GEMM shape / variant a few times (for ncu captures and quick timing): python tools/gemm_one.py V R N K [iters]"""
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from unigen_b200 import ops  # noqa: E402

v, R, N, K = (int(x) for x in sys.argv[1:5])
iters = int(sys.argv[5]) if len(sys.argv) > 5 else 20
torch.manual_seed(0)
a = torch.randn(1, R, K, device="cuda").to(torch.bfloat16)
w = (torch.randn(N, K, device="cuda") * K ** -0.5).to(torch.bfloat16)
out = torch.empty(1, R, N, device="cuda", dtype=torch.bfloat16)
for _ in range(3):
    ops.gemm(a, w, out=out, variant=v)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
    ops.gemm(a, w, out=out, variant=v)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
ref = (a.float() @ w.float().t())
rel = ((out.float() - ref).norm() / ref.norm()).item()
print(json.dumps({"variant": v, "shape": [R, N, K], "ms": ms, "tflops": 2.0 * R * N * K / ms / 1e9, "rel_l2": rel}))
