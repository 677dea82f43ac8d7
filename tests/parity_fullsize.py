"""Full-size parity of the B200-native UniGenFlux forward against the FP32 ORACLE run on the same B200.

At the BASELINE sizes (cfg2: 1024 + 1024 + 512 tokens, cfg3: 4096 + 4096 + 512 tokens; 19 + 38 base / 9 + 19 control blocks,
hidden 3072, 18.7 B parameters) the CPU oracle needs minutes per block, so the oracle restatement (oracle/unigen_oracle.py,
pure torch) runs in fp32 ON the GPU (TF32 off) over an fp32 view of the native model's own bf16 weights. Two checks:

  1. TEACHER-FORCED, per block: every block of the weave (base double / control double / zero-linear add / base single /
     control single / pre-stage pieces / embeddings / norm_out + proj_out) is evaluated by the oracle on the NATIVE trace's
     input of that block and compared with the native output of the same block: rel-L2 <= 1e-2 (north_star "per block").
     Routing is checked bit-exactly GIVEN the native bf16 gate input (argmax, RTS top-C, slots, counts) — a token whose two
     best fp32 logits are closer than 1e-5 may legally flip with the summation order; such tokens are counted, not failed.
  2. FREE-RUNNING: the fp32 oracle forward from the same inputs; final velocity cosine >= 0.999 (north_star) + rel-L2 and
     routing agreement reported.

python tests/parity_fullsize.py --workload cfg3 [--out gpurun_out/r02_parity_cfg3.json]     -> one JSON line
(test infrastructure: imports oracle/; never used by the product path; not collected by pytest — tests/test_fullsize_gpu.py
calls main())"""
import argparse
import json
import sys
import time
from pathlib import Path

import torch
import torch.nn.functional as F

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


class F32View:
    """fp32 view of a (bf16) state dict, upcast on access: the 18.7 B parameters never exist twice."""

    def __init__(self, sd):
        self.sd = sd

    def __getitem__(self, k):
        return self.sd[k].float()

    def get(self, k, default=None):
        v = self.sd.get(k)
        return default if v is None else v.float()

    def __contains__(self, k):
        return k in self.sd


def rel_l2(got, want):
    got, want = got.float(), want.float()
    return ((got - want).norm() / want.norm().clamp_min(1e-12)).item()


def routing_given_gate_input(O, trace, wg, rts_uniform, capacity, tag="moe"):
    """DeepSpeed top1gating on the NATIVE bf16 gate input, compared bit-exactly with the native routing."""
    G = trace[f"{tag}.gate_input"]
    logits = F.linear(G.reshape(-1, G.shape[-1]).float(), wg.float())
    _, _, _, counts, (idx, slot, _prob) = O.top1gating(logits, capacity, rts_uniform.float())
    n_idx, n_slot = trace[f"{tag}.route.expert_idx"].long(), trace[f"{tag}.route.slot"].long()
    top2 = logits.topk(2, dim=1).values
    near_tie = (top2[:, 0] - top2[:, 1]) <= 1e-5
    mism = idx != n_idx
    rec = dict(tokens=int(idx.numel()), near_tie_tokens=int(near_tie.sum()), expert_idx_mismatch=int(mism.sum()),
               mismatch_outside_near_ties=int((mism & ~near_tie).sum()))
    if rec["expert_idx_mismatch"] == 0:  # slots / counts depend on every token's expert: exact iff the argmax agrees everywhere
        st = torch.full((counts.numel() * capacity,), -1, dtype=torch.long, device=idx.device)
        kept = slot >= 0
        st[idx[kept] * capacity + slot[kept]] = torch.arange(idx.numel(), device=idx.device)[kept]
        rec.update(slot_mismatch=int((slot != n_slot).sum()),
                   slot_token_mismatch=int((st != trace[f"{tag}.route.slot_token"].long()).sum()),
                   exp_counts_equal=bool(torch.equal(counts, trace[f"{tag}.route.exp_counts"].long())),
                   dropped_tokens=int((slot < 0).sum()))
    rec["bit_exact"] = rec["mismatch_outside_near_ties"] == 0 and rec.get("slot_mismatch", 0) == 0 and \
        rec.get("slot_token_mismatch", 0) == 0 and rec.get("exp_counts_equal", True)
    return rec


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="cfg3", choices=["tiny", "cfg2", "cfg3"])
    ap.add_argument("--out", default=None)
    ap.add_argument("--no-free-run", action="store_true")
    args = ap.parse_args(argv)
    from oracle import unigen_oracle as O
    from unigen_b200 import ops
    from unigen_b200.model import FluxArch, UniGenFlux, canonical_control_params

    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    torch.set_float32_matmul_precision("highest")
    tiny = args.workload == "tiny"
    side = {"tiny": 256, "cfg2": 512, "cfg3": 1024}[args.workload]
    cfg = O.FluxConfig.tiny() if tiny else O.FluxConfig.flux()
    arch = FluxArch.tiny() if tiny else FluxArch()
    dev, bf = torch.device("cuda"), torch.bfloat16
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
    inp = dict(hidden_states=torch.randn(1, N, 64, device=dev, generator=g).to(bf),
               condition_hidden_states=torch.randn(1, N, 64, device=dev, generator=g).to(bf),
               encoder_hidden_states=torch.randn(1, T, 4096, device=dev, generator=g).to(bf),
               pooled_projections=torch.randn(1, 768, device=dev, generator=g),
               condition_pooled_projections=torch.randn(1, 768, device=dev, generator=g), timestep=torch.tensor([0.75], device=dev),
               img_ids=ids, txt_ids=torch.zeros(T, 3, device=dev), condition_ids=ids.clone(),
               rts_uniform=torch.rand(N, cfg.expert_nums, device=dev, generator=g))
    # ---- native forward with the per-block trace ----
    model.trace = {}
    ops.reset_launch_count()
    vel_native = model(**inp)[0].float().clone()
    torch.cuda.synchronize()
    launches = ops.launch_count()
    tr = model.trace
    model.trace = None

    sd = F32View(dict(model.state_dict()))
    H, D, E = cfg.num_attention_heads, cfg.inner_dim, cfg.expert_nums
    C = O.moe_capacity(N, E)
    scale = 1.0
    blocks = {}

    def cmp(name, got_native, want_oracle):
        blocks[name] = rel_l2(got_native, want_oracle)

    t0 = time.time()
    with torch.no_grad():
        hs32, cs32, es32 = (inp[k].float() for k in ("hidden_states", "condition_hidden_states", "encoder_hidden_states"))
        pooled, cpooled = inp["pooled_projections"], inp["condition_pooled_projections"]
        t1000 = inp["timestep"] * 1000
        # embeddings (src/UniGenTransformer.py:1215-1239)
        cmp("x_embed", tr["x_embed"], O.linear(sd, "x_embedder", hs32))
        cmp("context_embed", tr["context_embed"], O.linear(sd, "context_embedder", es32))
        cmp("temb", tr["temb"], O.combined_timestep_text_embed(sd, "time_text_embed", t1000, pooled))
        cmp("control_temb", tr["control_temb"], O.combined_timestep_text_embed(sd, "control_time_text_embed", t1000, pooled))
        cmp("condition_temb", tr["condition_temb"], O.combined_timestep_text_embed(sd, "control_condition_embed", t1000, cpooled))
        temb, ctemb, cdtemb = tr["temb"], tr["control_temb"], tr["condition_temb"]
        rope = O.flux_pos_embed(torch.cat([inp["txt_ids"], ids], 0), cfg.axes_dims_rope, cfg.theta)
        sched_d = O.weave_schedule(cfg.num_layers, cfg.cn_joint_layers)
        sched_s = O.weave_schedule(cfg.num_single_layers, cfg.cn_single_layers)

        h_in, c_in = tr["x_embed"], tr["context_embed"]
        routing = None
        for i in range(cfg.num_layers):
            c_o, h_o = O.flux_double_block(sd, f"transformer_blocks.{i}", H, h_in, c_in, temb, rope)
            cmp(f"double.{i}.base_hidden", tr[f"double.{i}.base_hidden"], h_o)
            cmp(f"double.{i}.base_context", tr[f"double.{i}.base_context"], c_o)
            base_h, base_c = tr[f"double.{i}.base_hidden"], tr[f"double.{i}.base_context"]
            j = sched_d[i]
            if i == 0:
                # ---- CoMoE pre-stage, teacher-forced (src/UniGenTransformer.py:1028-1068, 969-1026) ----
                cmp("moe.control_context", tr["moe.control_context"], O.linear(sd, "control_context_embedder", base_c))
                cmp("moe.cond_embed", tr["moe.cond_embed"], O.linear(sd, "control_x_embedder", cs32))
                gate_sum = (base_h.to(bf).float() + tr["moe.cond_embed"]).to(bf).float()
                blocks["moe.gate_input(bit-exact bf16 sum)"] = 0.0 if torch.equal(gate_sum, tr["moe.gate_input"]) else 1.0
                routing = routing_given_gate_input(O, tr, model.gate_wg, inp["rts_uniform"], C)
                # expert + shared-expert values under the NATIVE routing (a near-tie flip would change whole token rows)
                n_idx, n_slot, n_prob = (tr[f"moe.route.{k}"] for k in ("expert_idx", "slot", "prob"))
                real_top1 = O.top1gating

                def forced_top1(logits, capacity, rts):
                    S = logits.shape[0]
                    combine = torch.zeros(S, E, capacity, device=logits.device)
                    kept = n_slot >= 0
                    tok = torch.arange(S, device=logits.device)
                    combine[tok[kept], n_idx.long()[kept], n_slot.long()[kept]] = n_prob[kept]
                    return (torch.zeros((), device=logits.device), combine, combine.bool(), tr["moe.route.exp_counts"],
                            (n_idx.long(), n_slot.long(), n_prob))

                oracle = O.UniGenFluxOracle(cfg, sd)
                oracle.record = True
                O.top1gating = forced_top1
                try:
                    eh, ec, _, _ = oracle.moe_forward(base_h, tr["moe.cond_embed"], tr["moe.control_context"], ctemb, cdtemb, pooled,
                                                      cpooled, (ids, inp["txt_ids"], inp["condition_ids"]), inp["rts_uniform"])
                finally:
                    O.top1gating = real_top1
                for k in ("expert_hidden", "expert_cond", "shared_hidden", "shared_cond"):
                    cmp("moe." + k, tr["moe." + k], oracle.trace["moe." + k])
                cmp("moe.ctrl_in", tr["moe.ctrl_in"], eh + ec)
                ctrl_in = tr["moe.ctrl_in"]
                del oracle, eh, ec
            else:
                ctrl_in = base_h
            _, ch = O.flux_double_block(sd, f"control_joint_trans_blocks.{j}", H, ctrl_in, tr["moe.control_context"], cdtemb, rope)
            cmp(f"double.{i}.ctrl_hidden", tr[f"double.{i}.ctrl_hidden"], ch)
            add = base_h + O.linear(sd, f"controlnet_add_joint_blocks.{j}", tr[f"double.{i}.ctrl_hidden"]) * scale
            cmp(f"double.{i}.hidden", tr[f"double.{i}.hidden"], add)
            h_in, c_in = tr[f"double.{i}.hidden"], base_c
        x_in = torch.cat([c_in, h_in], 1)
        for i in range(cfg.num_single_layers):
            x_o = O.flux_single_block(sd, f"single_transformer_blocks.{i}", H, x_in, temb, rope)
            cmp(f"single.{i}.base_hidden", tr[f"single.{i}.base_hidden"], x_o)
            base_x = tr[f"single.{i}.base_hidden"]
            j = sched_s[i]
            cx = O.flux_single_block(sd, f"control_single_trans_blocks.{j}", H, base_x, cdtemb, rope)
            cmp(f"single.{i}.ctrl_hidden", tr[f"single.{i}.ctrl_hidden"], cx)
            add = base_x + O.linear(sd, f"controlnet_add_single_blocks.{j}", tr[f"single.{i}.ctrl_hidden"]) * scale
            cmp(f"single.{i}.hidden", tr[f"single.{i}.hidden"], add)
            x_in = tr[f"single.{i}.hidden"]
        out = O.linear(sd, "proj_out", O.ada_layer_norm_continuous(sd, "norm_out", x_in[:, T:], temb))
        cmp("velocity(norm_out+proj_out)", tr["velocity"], out)
    torch.cuda.synchronize()
    t_teacher = time.time() - t0
    worst = sorted(blocks.items(), key=lambda kv: -kv[1])[:8]
    rec = {"workload": args.workload, "tokens": {"image": N, "condition": N, "text": T}, "native_launches": launches,
           "oracle": "oracle/unigen_oracle.py in fp32 on the B200 (TF32 off) over an fp32 view of the native model's bf16 weights",
           "teacher_forced": {"blocks_checked": len(blocks), "max_rel_l2": max(blocks.values()),
                              "mean_rel_l2": sum(blocks.values()) / len(blocks), "bar": 1e-2,
                              "worst": [{"block": k, "rel_l2": v} for k, v in worst], "seconds": t_teacher},
           "routing_given_native_gate_input": routing, "per_block_rel_l2": blocks}
    del tr
    torch.cuda.empty_cache()
    if not args.no_free_run:
        t0 = time.time()
        oracle = O.UniGenFluxOracle(cfg, sd)
        oracle.record = False
        einp = {k: (v.float() if v.dtype == bf else v) for k, v in inp.items()}
        traced = {}
        oracle._rec = lambda name, t: traced.__setitem__(name, t.detach().clone()) if name == "moe.expert_idx" else None
        with torch.no_grad():
            vel_oracle = oracle.forward(**einp)[0].float()
        torch.cuda.synchronize()
        agree = (traced["moe.expert_idx"].long() == model._last_route["expert_idx"].long()).float().mean().item() \
            if "moe.expert_idx" in traced else None
        rec["free_running"] = {"cosine": F.cosine_similarity(vel_native.flatten(), vel_oracle.flatten(), dim=0).item(),
                               "rel_l2": rel_l2(vel_native, vel_oracle), "max_abs_err": (vel_native - vel_oracle).abs().max().item(),
                               "velocity_abs_max": vel_oracle.abs().max().item(), "routing_agreement": agree,
                               "bar": {"cosine": 0.999}, "seconds": time.time() - t0,
                               "note": "native bf16 end to end vs fp32 oracle end to end (57 + 28 blocks deep); the oracle routes on the "
                                       "unrounded fp32 sum, the native path (like the reference's bf16 run) on the bf16 sum"}
    line = json.dumps(rec)
    print(line)
    if args.out:
        Path(args.out).parent.mkdir(parents=True, exist_ok=True)
        Path(args.out).write_text(line + "\n")
    return rec


if __name__ == "__main__":
    main()
