"""GPU probe for the tcgen05 attention kernel: correctness against torch SDPA (fp32, same bf16 inputs, dense mask
expanded from the segment rule), then timing on the Flux-arch shapes. Output: gpurun_out/probe_attn.log"""
from __future__ import annotations

import ctypes as C
import json
import math
import os
import subprocess
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def run_variant(variant: int) -> None:
    import torch
    from unigen_b200 import _lib

    lib = _lib.load()
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    stream = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)  # noqa: E731

    def attn(qkv, H, dh, segs=None, vis=None):
        # qkv: [B, S, 3, H, dh] bf16 contiguous buffer (the layout the fused QKV GEMM writes)
        B, S = qkv.shape[:2]
        out = torch.zeros(B, S, H * dh, device=dev, dtype=torch.bfloat16)
        a = _lib.AttnArgs()
        rs, bs = qkv.stride(1), qkv.stride(0)
        a.q, a.k, a.v = qkv[:, :, 0].data_ptr(), qkv[:, :, 1].data_ptr(), qkv[:, :, 2].data_ptr()
        a.o = out.data_ptr()
        a.q_row_stride = a.k_row_stride = a.v_row_stride = rs
        a.q_batch_stride = a.k_batch_stride = a.v_batch_stride = bs
        a.o_row_stride, a.o_batch_stride = out.stride(1), out.stride(0)
        a.batch, a.heads, a.seq, a.head_dim = B, H, S, dh
        a.scale = 1.0 / math.sqrt(dh)
        keep = []
        if segs is not None:
            n = len(segs) - 1
            sb = (C.c_int32 * (n + 1))(*segs)
            sv = (C.c_uint32 * n)(*vis)
            keep = [sb, sv]
            a.n_seg, a.seg_bounds, a.seg_visible = n, sb, sv
        a.variant = variant
        _lib.check(lib.ug_attention_bf16(C.byref(a), stream()), "attention")
        return out

    def dense_mask(S, segs, vis):
        n = len(segs) - 1
        seg_of = torch.zeros(S, dtype=torch.long)
        for i in range(n):
            seg_of[segs[i]:segs[i + 1]] = i
        vm = torch.tensor([[(vis[i] >> j) & 1 for j in range(n)] for i in range(n)], dtype=torch.bool)
        return vm[seg_of][:, seg_of].to(dev)

    def ref(qkv, H, dh, mask=None):
        q, k, v = (qkv[:, :, i].float().transpose(1, 2) for i in range(3))  # [B,H,S,dh]
        o = torch.nn.functional.scaled_dot_product_attention(q, k, v, attn_mask=mask)
        return o.transpose(1, 2).reshape(qkv.shape[0], qkv.shape[1], H * dh)

    def report(name, got, want, extra=None):
        diff = got.float() - want
        rel = (diff.norm() / want.norm().clamp_min(1e-12)).item()
        rec = {"variant": variant, "case": name, "rel_l2": rel, "max_abs": diff.abs().max().item(), "ok": bool(rel < 1e-2)}
        if not rec["ok"]:
            d = diff.abs()
            rowerr = d.reshape(-1, d.shape[-1]).max(1).values
            colerr = d.reshape(-1, d.shape[-1]).max(0).values
            thr = 0.05 * want.abs().max()
            rec["bad_rows_first"] = [int(i) for i in (rowerr > thr).nonzero().flatten()[:24]]
            rec["n_bad_rows"] = int((rowerr > thr).sum())
            rec["bad_cols_first"] = [int(i) for i in (colerr > thr).nonzero().flatten()[:24]]
            rec["n_bad_cols"] = int((colerr > thr).sum())
            rec["got"] = got.float().reshape(-1, got.shape[-1])[0, :6].tolist()
            rec["want"] = want.reshape(-1, want.shape[-1])[0, :6].tolist()
            rec["nan"] = bool(torch.isnan(got.float()).any())
        if extra:
            rec.update(extra)
        print(json.dumps(rec), flush=True)
        return rec["ok"]

    def mk(B, S, H, dh, scale=1.0):
        return (torch.randn(B, S, 3, H, dh, device=dev) * scale).to(torch.bfloat16)

    ok = True
    cases = [("one_tile", 1, 128, 1, 128), ("two_tiles", 1, 256, 2, 128), ("four_tiles", 1, 512, 2, 128),
             ("tail", 1, 300, 3, 128), ("dh64", 1, 768, 6, 64), ("dh64_tail", 2, 333, 2, 64),
             ("batch2", 2, 640, 4, 128), ("flux_small", 1, 1536, 24, 128), ("big_logits", 1, 512, 2, 128)]
    for name, B, S, H, dh in cases:
        qkv = mk(B, S, H, dh, scale=3.0 if name == "big_logits" else 1.0)
        try:
            out = attn(qkv, H, dh)
            torch.cuda.synchronize()
        except Exception as e:  # noqa: BLE001
            print(json.dumps({"variant": variant, "case": name, "error": str(e)}), flush=True)
            return
        ok &= report(name, out, ref(qkv, H, dh), {"shape": [B, S, H, dh]})

    # segment masks: [txt | img | c1 | c2]; P-variant rule (txt,img see all; c_i sees txt,img,c_i) and the
    # north-star's stricter rule (condition tokens see only themselves)
    for name, segs, vis in [
        ("seg_pvariant", [0, 64, 264, 464, 600], [0b1111, 0b1111, 0b0111, 0b1011]),
        ("seg_strict", [0, 64, 264, 464, 600], [0b1111, 0b1111, 0b0100, 0b1000]),
        ("seg_aligned", [0, 128, 384, 640], [0b111, 0b111, 0b100]),
        ("seg_blockdiag", [0, 100, 356, 612], [0b001, 0b010, 0b100]),
    ]:
        S = segs[-1]
        qkv = mk(2, S, 3, 128)
        try:
            out = attn(qkv, 3, 128, segs, vis)
            torch.cuda.synchronize()
        except Exception as e:  # noqa: BLE001
            print(json.dumps({"variant": variant, "case": name, "error": str(e)}), flush=True)
            return
        ok &= report(name, out, ref(qkv, 3, 128, dense_mask(S, segs, vis)))
        m = torch.zeros(S, S, dtype=torch.uint8, device=dev)
        sb = (C.c_int32 * len(segs))(*segs)
        sv = (C.c_uint32 * len(vis))(*vis)
        _lib.check(lib.ug_expand_segment_mask(S, len(vis), sb, sv, m.data_ptr(), stream()))
        torch.cuda.synchronize()
        print(json.dumps({"variant": variant, "case": name + "_mask_bitexact",
                          "ok": bool(torch.equal(m.bool(), dense_mask(S, segs, vis)))}), flush=True)
    if not ok:
        return
    for (B, S, H, dh) in [(1, 4608, 24, 128), (1, 8704, 24, 128), (1, 1536, 24, 128), (2, 4608, 24, 128), (2, 4429, 24, 64),
                          (1, 4608, 3, 128), (1, 16896, 24, 128)]:
        qkv = mk(B, S, H, dh)
        for _ in range(3):
            attn(qkv, H, dh)
        torch.cuda.synchronize()
        iters = 10
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            attn(qkv, H, dh)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        q, k, v = (qkv[:, :, i].transpose(1, 2) for i in range(3))
        for _ in range(3):
            torch.nn.functional.scaled_dot_product_attention(q, k, v)
        e0.record()
        for _ in range(iters):
            torch.nn.functional.scaled_dot_product_attention(q, k, v)
        e1.record()
        torch.cuda.synchronize()
        ms_ref = e0.elapsed_time(e1) / iters
        fl = 4.0 * B * H * S * S * dh
        print(json.dumps({"variant": variant, "case": "timing", "shape": [B, S, H, dh], "ms": ms,
                          "tflops": fl / ms / 1e9, "sdpa_ms": ms_ref, "sdpa_tflops": fl / ms_ref / 1e9}), flush=True)


def main() -> None:
    if len(sys.argv) >= 3 and sys.argv[1] == "--variant":
        run_variant(int(sys.argv[2]))
        return
    out_dir = ROOT / "gpurun_out"
    out_dir.mkdir(exist_ok=True)
    log = open(out_dir / "probe_attn.log", "w")
    for v in [int(x) for x in os.environ.get("UG_PROBE_VARIANTS", "3,1").split(",")]:
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, __file__, "--variant", str(v)], capture_output=True, text=True, timeout=300)
            log.write(r.stdout)
            log.write(f"# variant {v} exit {r.returncode} in {time.time() - t0:.1f}s\n")
            if r.returncode != 0 or "error" in r.stdout:
                log.write("# stderr tail:\n" + r.stderr[-3000:] + "\n")
        except subprocess.TimeoutExpired as e:
            log.write(str(e.stdout or ""))
            log.write(f"# variant {v} TIMEOUT\n")
        log.flush()
    log.close()
    print(open(out_dir / "probe_attn.log").read())


if __name__ == "__main__":
    main()
