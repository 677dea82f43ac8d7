import sys, torch
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import test_sd3_gpu as t
for seed in (5, 6, 7):
    cfg, sd, inp, oracle, model = t._setup(height=320, width=192, text_len=333, batch=2, seed=seed)
    inp["conditioning_scale"] = 0.6
    want, _, want_o = oracle.forward(**inp)
    model.trace = {}
    got, _, outs = model(**t._dev(inp))
    idx = model._last_route["expert_idx"].cpu().long(); slot = model._last_route["slot"].cpu().long()
    print("seed", seed, "idx agree", (idx == oracle.trace["moe.expert_idx"]).float().mean().item(), "slot agree", (slot == oracle.trace["moe.slot"]).float().mean().item(),
          "counts", outs["expert_counts"].tolist(), want_o["expert_counts"].tolist())
    for k in ("moe.cond_embed","moe.enc_ctrl","block.0.base_hidden","moe.expert_hidden","moe.expert_cond","moe.shared_hidden","moe.shared_cond","moe.ctrl_in","velocity"):
        print("   ", k, t.rel_l2(model.trace[k], oracle.trace[k]))
    # per-token error of expert_hidden for tokens whose routing agrees
    ok = (idx == oracle.trace["moe.expert_idx"]) & (slot == oracle.trace["moe.slot"])
    eh, ehw = model.trace["moe.expert_hidden"].cpu().reshape(-1, cfg.inner_dim), oracle.trace["moe.expert_hidden"].reshape(-1, cfg.inner_dim)
    err = (eh - ehw).norm(dim=1) / ehw.norm(dim=1).clamp_min(1e-6)
    print("    per-token err (routing-agree tokens) max", err[ok & (slot >= 0)].max().item(), "mean", err[ok & (slot>=0)].mean().item(), "n_disagree", (~ok).sum().item())
