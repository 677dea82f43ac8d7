"""GPU probe for the tcgen05 GEMM: correctness against torch fp32 matmul on the same bf16 inputs, then timing.
Run on a B200: `python tools/probe_gemm.py` (spawns one subprocess per kernel variant so a trapped launch in one
variant cannot poison the others). Output: gpurun_out/probe_gemm.log"""
from __future__ import annotations

import ctypes as C
import json
import os
import subprocess
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def run_variant(variant: int, quick: bool) -> None:
    import torch
    from unigen_b200._lib import GemmArgs, LIB_PATH

    lib = C.CDLL(str(LIB_PATH))
    lib.ug_last_error.restype = C.c_char_p
    lib.ug_gemm_bf16.argtypes = [C.POINTER(GemmArgs), C.c_void_p]
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    print(json.dumps({"variant": variant, "device_check": lib.ug_device_check(),
                      "msg": lib.ug_last_error().decode()}), flush=True)

    def gemm(a, w, bias=None, gate=None, alpha=1.0, act=0, res=None, out=None):
        # a: [B, R, K] (possibly strided view), w: [N, K] or [B, N, K]
        B, R, K = a.shape
        N = w.shape[-2]
        if out is None:
            out = torch.empty(B, R, N, device=dev, dtype=torch.bfloat16)
        g = GemmArgs()
        g.a, g.a_row_stride, g.a_batch_stride = a.data_ptr(), a.stride(1), a.stride(0)
        g.w, g.w_row_stride = w.data_ptr(), w.stride(-2)
        g.w_batch_stride = w.stride(0) if w.dim() == 3 else 0
        g.c, g.c_row_stride, g.c_batch_stride = out.data_ptr(), out.stride(1), out.stride(0)
        g.batch, g.rows, g.n, g.k = B, R, N, K
        if bias is not None:
            g.bias = bias.data_ptr()
            g.bias_batch_stride = bias.stride(0) if bias.dim() == 2 else 0
        if gate is not None:
            g.gate, g.gate_batch_stride = gate.data_ptr(), gate.stride(0)
        g.alpha, g.act = alpha, act
        if res is not None:
            g.residual, g.res_row_stride, g.res_batch_stride = res.data_ptr(), res.stride(1), res.stride(0)
        g.variant = variant
        st = lib.ug_gemm_bf16(C.byref(g), C.c_void_p(torch.cuda.current_stream().cuda_stream))
        if st != 0:
            raise RuntimeError(f"ug_gemm_bf16 -> {st}: {lib.ug_last_error().decode()}")
        return out

    def ref(a, w, bias=None, gate=None, alpha=1.0, act=0, res=None):
        y = torch.matmul(a.float(), w.float().transpose(-1, -2))
        if bias is not None:
            y = y + (bias.float()[:, None, :] if bias.dim() == 2 else bias.float())
        if act == 1:
            y = torch.nn.functional.gelu(y, approximate="tanh")
        if gate is not None:
            y = y * gate[:, None, :] * alpha
        elif alpha != 1.0:
            y = y * alpha
        if res is not None:
            y = y + res.float()
        return y

    def report(name, got, want, extra=None):
        got = got.float()
        diff = (got - want)
        rel = (diff.norm() / want.norm().clamp_min(1e-12)).item()
        rec = {"variant": variant, "case": name, "rel_l2": rel, "max_abs": diff.abs().max().item(),
               "ok": bool(rel < 6e-3)}
        if not rec["ok"]:
            # structure of the error: which 8-column groups / 8-row groups are wrong (swizzle / layout diagnosis)
            d2 = diff.reshape(-1, diff.shape[-1]).abs()
            colerr = d2.max(0).values
            rowerr = d2.max(1).values
            rec["bad_cols_first64"] = [int(i) for i in (colerr[:64] > 0.05 * want.abs().max()).nonzero().flatten()[:64]]
            rec["bad_rows_first64"] = [int(i) for i in (rowerr[:64] > 0.05 * want.abs().max()).nonzero().flatten()[:64]]
            rec["got_sample"] = got.reshape(-1, got.shape[-1])[0, :8].tolist()
            rec["want_sample"] = want.reshape(-1, want.shape[-1])[0, :8].tolist()
        if extra:
            rec.update(extra)
        print(json.dumps(rec), flush=True)
        return rec["ok"]

    def rnd(*shape, scale=1.0):
        return (torch.randn(*shape, device=dev) * scale).to(torch.bfloat16)

    bm = 256 if variant == 2 else 128
    bn = 128 if variant == 3 else 256
    cases = [
        ("one_tile_k64", 1, bm, bn, 64),
        ("one_tile_k256", 1, bm, bn, 256),
        ("multi_tile", 1, 2 * bm, 2 * bn, 512),
        ("ragged", 1, 200, 384, 192),
        ("ragged2", 1, 1000, 1152, 384),
        ("narrow_n64", 1, 300, 64, 384),
        ("k_tail", 1, 256, 256, 72),
        ("many_tiles", 1, 4608, 3072, 1024),
    ]
    all_ok = True
    for name, B, R, N, K in cases:
        a, w = rnd(B, R, K), rnd(N, K, scale=K ** -0.5)
        try:
            out = gemm(a, w)
            torch.cuda.synchronize()
        except Exception as e:  # noqa: BLE001
            print(json.dumps({"variant": variant, "case": name, "error": str(e)}), flush=True)
            return
        all_ok &= report(name, out, ref(a, w), {"shape": [B, R, N, K]})

    # fused epilogue + batched / strided views
    try:
        B, R, N, K = 3, 333, 640, 256
        big = rnd(B, R + 50, K + 64)
        a = big[:, 50:, 64:]  # strided view (row stride K+64, offset)
        w = rnd(N, K, scale=K ** -0.5)
        bias = rnd(N)
        gate = torch.randn(B, N, device=dev)
        res_full = rnd(B, R, N + 128)
        res = res_full[:, :, 128:]
        out_full = torch.zeros(B, R + 7, N + 64, device=dev, dtype=torch.bfloat16)
        out = out_full[:, 7:, 64:]
        gemm(a, w, bias=bias, gate=gate, alpha=0.5, act=1, res=res, out=out)
        torch.cuda.synchronize()
        all_ok &= report("epilogue_strided_batched", out, ref(a, w, bias, gate, 0.5, 1, res))
        untouched = float(out_full[:, :7].abs().max()) + float(out_full[:, :, :64].abs().max())
        print(json.dumps({"variant": variant, "case": "no_oob_writes", "ok": untouched == 0.0}), flush=True)
        # batched weights (expert-style) + per-batch bias
        wb = rnd(B, N, K, scale=K ** -0.5)
        bb = rnd(B, N)
        out = gemm(a, wb, bias=bb)
        torch.cuda.synchronize()
        all_ok &= report("batched_weights", out, ref(a, wb, bb))
        # in-place residual (c aliases residual)
        h = rnd(B, R, N)
        h0 = h.clone()
        gemm(a, w, bias=bias, gate=gate, res=h, out=h)
        torch.cuda.synchronize()
        all_ok &= report("inplace_residual", h, ref(a, w, bias, gate, 1.0, 0, h0))
    except Exception as e:  # noqa: BLE001
        print(json.dumps({"variant": variant, "case": "epilogue", "error": str(e)}), flush=True)
        return

    if quick or not all_ok:
        return
    # timing: the Flux-arch shapes of cfg3 (S = 4608)
    for (R, N, K) in [(4608, 3072, 3072), (4608, 9216, 3072), (4608, 12288, 3072), (4608, 3072, 12288),
                      (4608, 3072, 15360), (4096, 3072, 3072), (512, 3072, 3072), (8192, 8192, 8192)]:
        a, w = rnd(1, R, K), rnd(N, K, scale=K ** -0.5)
        out = torch.empty(1, R, N, device=dev, dtype=torch.bfloat16)
        for _ in range(3):
            gemm(a, w, out=out)
        torch.cuda.synchronize()
        iters = 20
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            gemm(a, w, out=out)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        # cuBLAS comparator (context only)
        a2, w2 = a[0], w
        for _ in range(3):
            torch.matmul(a2, w2.t())
        e0.record()
        for _ in range(iters):
            torch.matmul(a2, w2.t())
        e1.record()
        torch.cuda.synchronize()
        ms_ref = e0.elapsed_time(e1) / iters
        fl = 2.0 * R * N * K
        print(json.dumps({"variant": variant, "case": "timing", "shape": [R, N, K], "ms": ms,
                          "tflops": fl / ms / 1e9, "cublas_ms": ms_ref, "cublas_tflops": fl / ms_ref / 1e9}),
              flush=True)


def main() -> None:
    if len(sys.argv) >= 3 and sys.argv[1] == "--variant":
        run_variant(int(sys.argv[2]), quick="--quick" in sys.argv)
        return
    out_dir = ROOT / "gpurun_out"
    out_dir.mkdir(exist_ok=True)
    log = open(out_dir / "probe_gemm.log", "w")
    variants = [int(v) for v in os.environ.get("UG_PROBE_VARIANTS", "1,3,2").split(",")]
    for v in variants:
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, __file__, "--variant", str(v)], capture_output=True, text=True,
                               timeout=300)
            log.write(r.stdout)
            log.write(f"# variant {v} exit {r.returncode} in {time.time() - t0:.1f}s\n")
            if r.returncode != 0 or "error" in r.stdout:
                log.write("# stderr tail:\n" + r.stderr[-3000:] + "\n")
        except subprocess.TimeoutExpired as e:
            log.write((e.stdout or b"").decode() if isinstance(e.stdout, bytes) else (e.stdout or ""))
            log.write(f"# variant {v} TIMEOUT\n")
        log.flush()
    log.close()
    print(open(out_dir / "probe_gemm.log").read())


if __name__ == "__main__":
    main()
