"""VAE encode / decode timing on one B200 (SURVEY.md §8 (f)4 tail): FLUX.1 VAE architecture, random-init weights, 1024 x 1024.

Prints ONE JSON line: ms per decode / encode (CUDA events, L2-exceeding working set), the convolution GEMMs' achieved TFLOP/s from an
instrumented pass (CUDA events around every conv / GEMM launch: implicit 3x3 vs patch-gather + GEMM vs plain GEMM), the HBM-bound
kernels' GB/s, and — optionally — the same network as torch-eager bf16 (cuDNN) on the same GPU.
    python tools/bench_vae.py [--side 1024] [--batch 1] [--steps 5] [--eager]
"""
from __future__ import annotations

import argparse
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--side", type=int, default=1024)
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--eager", action="store_true", help="also time the oracle network as torch-eager bf16 (cuDNN) on the same GPU")
    args = ap.parse_args()
    from unigen_b200 import ops
    from unigen_b200.vae import AutoencoderKL
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    vae = AutoencoderKL(device=dev).init_random_(seed=0)
    g = torch.Generator(device=dev).manual_seed(1)
    B, S = args.batch, args.side
    lat = torch.randn(B, 16, S // 8, S // 8, device=dev, generator=g).to(torch.bfloat16)
    img = (torch.rand(B, 3, S, S, device=dev, generator=g) * 2 - 1).to(torch.bfloat16)

    def timed(fn, n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    dec = lambda: vae.decode_latents(lat)  # noqa: E731
    enc = lambda: vae.encode_condition(img, sample=False)  # noqa: E731
    for _ in range(args.warmup):
        dec(), enc()
    n0 = ops.launch_count()
    ms_dec = timed(dec, args.steps)
    launches_dec = (ops.launch_count() - n0) // args.steps
    ms_enc = timed(enc, args.steps)

    # ---- instrumented pass: per-launch CUDA events by op class ----
    recs = {}
    orig = {k: getattr(ops, k) for k in ("conv3x3", "gemm", "im2col", "groupnorm", "upsample2x", "softmax_rows_")}

    def wrap(kind, fn, work):
        def f(*a, **k):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r = fn(*a, **k)
            e1.record()
            recs.setdefault(kind, []).append((e0, e1, float(work(r, *a, **k))))
            return r
        return f

    def conv_flops(_r, x, w, *a, **k):
        return 2.0 * x.shape[0] * x.shape[1] * x.shape[2] * w.shape[0] * w.shape[1]

    def gemm_flops(_r, a, w, *p, **k):
        a3 = a if a.dim() == 3 else a.unsqueeze(0)
        return 2.0 * a3.shape[0] * a3.shape[1] * w.shape[-2] * w.shape[-1]

    ops.conv3x3 = wrap("conv3x3_implicit", orig["conv3x3"], conv_flops)
    ops.gemm = wrap("gemm", orig["gemm"], gemm_flops)
    ops.im2col = wrap("im2col", orig["im2col"], lambda _r, x, *a, **k: 2.0 * _r.numel())  # bytes written (reads <= that)
    ops.groupnorm = wrap("groupnorm", orig["groupnorm"], lambda _r, x, *a, **k: 6.0 * x.numel())  # 2 reads + 1 write, bf16
    ops.upsample2x = wrap("upsample2x", orig["upsample2x"], lambda _r, x, *a, **k: 2.0 * (x.numel() + _r.numel()))
    ops.softmax_rows_ = wrap("softmax_rows", orig["softmax_rows_"], lambda _r, x, *a, **k: 8.0 * x.numel())  # 3 reads + 1 write
    try:
        dec()
        torch.cuda.synchronize()
        dec_recs, recs = recs, {}
        enc()
        torch.cuda.synchronize()
        enc_recs = recs
    finally:
        for k, v in orig.items():
            setattr(ops, k, v)

    def summarise(rs):
        out = {}
        for kind, lst in rs.items():
            ms = sum(e0.elapsed_time(e1) for e0, e1, _ in lst)
            work = sum(w for _, _, w in lst)
            unit = "TFLOP/s" if kind in ("conv3x3_implicit", "gemm") else "GB/s"
            rate = work / ms / (1e9 if unit == "TFLOP/s" else 1e6) if ms > 0 else 0.0
            out[kind] = {"launches": len(lst), "ms": ms, "achieved": rate, "unit": unit, ("flops" if unit == "TFLOP/s" else "bytes"): work}
        return out

    d, e = summarise(dec_recs), summarise(enc_recs)
    peaks = {}
    try:
        peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
    except Exception:  # noqa: BLE001
        pass
    tf_dec = sum(v["flops"] for v in d.values() if "flops" in v) / 1e12
    tf_enc = sum(v["flops"] for v in e.values() if "flops" in v) / 1e12
    line = {"metric": "VAE decode / encode latency (FLUX.1 VAE architecture, random init)", "side": S, "batch": B, "dtype": "bf16",
            "decode_ms": ms_dec, "encode_ms": ms_enc, "decode_tflop": tf_dec, "encode_tflop": tf_enc,
            "decode_tflops": tf_dec / (ms_dec / 1e3), "encode_tflops": tf_enc / (ms_enc / 1e3), "decode_launches": launches_dec,
            "decode_ops": d, "encode_ops": e,
            "peaks": {"bf16_tflops_sustained": peaks.get("bf16_tflops_sustained"), "hbm_gbs": peaks.get("hbm_gbs")},
            "method": "ms = CUDA events over back-to-back calls; *_ops = one extra pass with an event pair around every launch (includes "
                      "the ~5 us event floor per launch)"}
    if args.eager:
        from oracle import vae_oracle as O
        cfg = O.VAEConfig.flux()
        sd = {k: v.detach().to(torch.bfloat16) for k, v in vae.state_dict().items()}
        orc = O.VAEOracle(cfg, sd)
        with torch.no_grad():
            ed = lambda: orc.decode(lat)  # noqa: E731
            ee = lambda: orc.encode(img, None)  # noqa: E731
            for _ in range(2):
                ed(), ee()
            line["eager_bf16"] = {"decode_ms": timed(ed, 3), "encode_ms": timed(ee, 3), "stack": f"torch {torch.__version__} eager bf16 (cuDNN conv2d, F.group_norm)"}
            got = vae.decode_latents(lat).float()
            want = ed().float()
            line["eager_bf16"]["decode_cosine_vs_native"] = torch.nn.functional.cosine_similarity(got.flatten(), want.flatten(), dim=0).item()
    print(json.dumps(line))


if __name__ == "__main__":
    main()
