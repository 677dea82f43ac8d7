"""One launch of every hot kernel at its cfg3 shape inside a cudaProfilerStart/Stop window, for ONE `ncu --set full` capture:

  ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/prof_r02_targets \\
      python tools/ncu_targets.py [--workload cfg3|cfg5]

cfg3 (Flux, D = 3072, dh = 128, S = 4608):  grouped AdaLN GEMV (the step's `late` job table, ~9 GB of weights), LN-modulate
(4608 x 3072), q|k|v projection GEMM (N = 9216) with the plain and with the fused QK-norm epilogue, the separate QK-RMSNorm + RoPE pass (in place, 4608 x 6144), joint attention (24 heads),
proj_mlp GEMM with the GELU epilogue (N = 12288), ff2-shaped GEMM with the gated-residual epilogue (K = 12288), proj_out of the
single blocks (K = 15360). cfg5 (SD3.5-medium, D = 1536, dh = 64): attention at dh = 64 and the K = 1536 GEMMs.
The model is built exactly as bench.py builds it, one warm-up forward allocates the workspace and job tables, and the window
then launches the ops on the workspace buffers (ncu flushes caches between replays: these are cold-cache numbers; the in-situ
bandwidth of the HBM-bound kernels is in bench.py's `roofline.hbm_bound_kernels`)."""
import argparse
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="cfg3", choices=["cfg3", "cfg5", "tiny", "vae"])
    args = ap.parse_args()
    from unigen_b200 import ops
    from unigen_b200.ops import UG_ACT_GELU_TANH
    dev = torch.device("cuda")
    prof = torch.cuda.profiler
    if args.workload == "vae":
        # FLUX.1 VAE decoder at 1024 x 1024, top level (128 channels) and the level below (256 channels at 512 x 512): the implicit
        # 3x3 convolution, GroupNorm + SiLU (statistics, fold, apply), nearest x2 up-sampling, the mid attention's row softmax
        bf = torch.bfloat16
        g = torch.Generator(device=dev).manual_seed(0)
        x1 = torch.randn(1, 1024, 1024, 128, device=dev, generator=g).to(bf)
        x2 = torch.randn(1, 512, 512, 256, device=dev, generator=g).to(bf)
        w1 = (torch.randn(128, 9 * 128, device=dev, generator=g) / 34).to(bf)
        w2 = (torch.randn(256, 9 * 256, device=dev, generator=g) / 48).to(bf)
        b1, b2 = torch.zeros(128, device=dev, dtype=bf), torch.zeros(256, device=dev, dtype=bf)
        gm, bt = torch.ones(128, device=dev, dtype=bf), torch.zeros(128, device=dev, dtype=bf)
        y1, y2, n1 = torch.empty_like(x1), torch.empty_like(x2), torch.empty_like(x1)
        up = torch.empty(1, 1024, 1024, 256, device=dev, dtype=bf)
        s = torch.randn(16384, 16384, device=dev, generator=g).to(bf)
        for _ in range(2):
            ops.conv3x3(x1, w1, bias=b1, residual=x1, out=y1)
        torch.cuda.synchronize()
        prof.start()
        ops.conv3x3(x1, w1, bias=b1, residual=x1, out=y1)
        ops.conv3x3(x2, w2, bias=b2, out=y2)
        ops.groupnorm(x1, gm, bt, 32, silu_act=True, out=n1)
        ops.upsample2x(x2, out=up)
        ops.softmax_rows_(s)
        torch.cuda.synchronize()
        prof.stop()
        return
    if args.workload == "cfg5":
        from unigen_b200.sd3 import SD3Arch, UniGenSD3, shipped_control_params
        model = UniGenSD3(SD3Arch(), device=dev)
        model.init_condition_block(condition_nums=1, control_params=shipped_control_params())
        model.init_random_(seed=0)
        B, lat, T = 2, 128, 333
        N = (lat // 2) ** 2
        g = torch.Generator(device=dev).manual_seed(1234)
        bf = torch.bfloat16
        inp = dict(hidden_states=torch.randn(B, 16, lat, lat, device=dev, generator=g).to(bf),
                   condition_hidden_states=torch.randn(B, 16, lat, lat, device=dev, generator=g).to(bf),
                   encoder_hidden_states=torch.randn(B, T, 4096, device=dev, generator=g).to(bf),
                   pooled_projections=torch.randn(B, 2048, device=dev, generator=g),
                   condition_pooled_projections=torch.randn(B, 2048, device=dev, generator=g),
                   timestep=torch.full((B,), 500.0, device=dev))
        model(**inp)
        torch.cuda.synchronize()
        buf, D, S = model._buf, model.inner_dim, N + T
        w = model.blocks[0]
        a = model.arch
        prof.start()
        ops.gemm(buf.NX[:, :S], w.attn.qkv[0], out=buf.QKV[:, :S], bias=w.attn.qkv[1])
        ops.attention(buf.QKV[:, :S, 0:D], buf.QKV[:, :S, D:2 * D], buf.QKV[:, :S, 2 * D:3 * D], buf.AO[:, :S], a.num_attention_heads,
                      a.attention_head_dim)
        torch.cuda.synchronize()
        prof.stop()
        return
    from unigen_b200.model import FluxArch, UniGenFlux, canonical_control_params
    tiny = args.workload == "tiny"
    arch = FluxArch.tiny() if tiny else FluxArch()
    side = 256 if tiny else 1024
    model = UniGenFlux(arch, device=dev)
    model.init_condition_block(condition_nums=1, control_params=canonical_control_params())
    model.init_random_(seed=0)
    T, grid = 512, side // 16
    N = grid * grid
    g = torch.Generator(device=dev).manual_seed(1234)
    ids = torch.zeros(grid, grid, 3, device=dev)
    ids[..., 1] += torch.arange(grid, device=dev)[:, None]
    ids[..., 2] += torch.arange(grid, device=dev)[None, :]
    ids = ids.reshape(N, 3)
    bf = torch.bfloat16
    inp = dict(hidden_states=torch.randn(1, N, 64, device=dev, generator=g).to(bf),
               condition_hidden_states=torch.randn(1, N, 64, device=dev, generator=g).to(bf),
               encoder_hidden_states=torch.randn(1, T, 4096, device=dev, generator=g).to(bf),
               pooled_projections=torch.randn(1, 768, device=dev, generator=g),
               condition_pooled_projections=torch.randn(1, 768, device=dev, generator=g), timestep=torch.tensor([0.75], device=dev),
               img_ids=ids, txt_ids=torch.zeros(T, 3, device=dev), condition_ids=ids.clone())
    model(**inp)
    torch.cuda.synchronize()
    buf, D, S = model._buf, model.inner_dim, T + N
    H, dh = arch.num_attention_heads, arch.attention_head_dim
    w = model.single[0]
    mp = model._mod_plans(buf)
    shift, scale, gate = mp.m_single[0]
    x, nx, cat = buf.X, buf.NX[:, :S], buf.CAT[:, :S]
    prof.start()
    ops.gemv_grouped(mp.late)                                                                   # AdaLN weights, HBM-bound
    ops.ln_modulate(x, nx, shift, scale)                                                        # HBM-bound
    ops.gemm(nx, w.qkv[0], out=buf.QKV[:, :S], bias=w.qkv[1])                                   # N = 3D, plain epilogue
    ops.gemm(nx, w.qkv[0], out=buf.QKV[:, :S], bias=w.qkv[1],                                   # the default: QK-norm + RoPE fused
             qk_norm=dict(weight=w.rms, head_dim=dh, d=D, cos_sin=buf.rope[:S], eps=1e-6))
    ops.gemm(nx, w.qkv[0], out=buf.QKV[:, :S], bias=w.qkv[1])                                   # (restore un-normalised q | k)
    ops.qk_rmsnorm_rope(buf.QKV[:, :S, :2 * D], 2 * H, dh, w.rms, buf.rope[:S], heads_per_weight=H)  # HBM-bound, in place
    ops.attention(buf.QKV[:, :S, 0:D], buf.QKV[:, :S, D:2 * D], buf.QKV[:, :S, 2 * D:3 * D], cat[:, :, :D], H, dh)
    ops.gemm(nx, w.mlp[0], out=cat[:, :, D:], bias=w.mlp[1], act=UG_ACT_GELU_TANH)              # N = 4D, GELU epilogue
    ops.gemm(cat, w.out[0], out=buf.CS, bias=w.out[1], gate=gate, residual=x)                   # K = 5D, gated residual
    dw = model.double[0]
    ops.gemm(buf.FF[:, :N], dw.ff2[0], out=buf.CH, bias=dw.ff2[1], gate=gate, residual=buf.CH)  # K = 4D, in-place residual
    torch.cuda.synchronize()
    prof.stop()


if __name__ == "__main__":
    main()
