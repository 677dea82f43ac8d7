"""Turn the raw artefacts a GPU call left under gpurun_out/ into the committed evidence under profiles/ (round 2: `r02_*`).

  python tools/make_profiles.py <tag>            e.g. tag = r02a: reads gpurun_out/<tag>_*.json / .csv and gpurun_out/prof_r02_targets*.ncu-rep

Writes
  profiles/r02_bench_*.json              copies of the bench lines
  profiles/r02_launches_cfg3.csv + _summary.txt   the ncu launch list (gpu__time_duration.sum per launch) and its per-kernel shares
  profiles/r02_ncu_targets_summary.txt   one row per hot kernel of the `ncu --set full` capture: duration, DRAM bytes read / written,
                                         DRAM throughput %, tensor-pipe / XU %, L2 hit rate, registers, grid
  profiles/r02_ncu_traffic.json          dram bytes of the dominant GEMM launch (bench.py's roofline.traffic reads this file)
  profiles/r02_sass_opcodes.txt          cuobjdump -sass opcode histogram per kernel of the in-tree libunigen_b200.so
"""
import collections
import csv
import json
import re
import shutil
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
OUT, PROF = ROOT / "gpurun_out", ROOT / "profiles"

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
           "lts__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
           "launch__grid_size", "launch__block_size", "sm__cycles_elapsed.avg.per_second"]


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1)


def to_us(v, unit):
    v = float(v.replace(",", ""))
    return v * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(unit, 1)


def ncu_rows(rep):
    raw = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    out = []
    for r in rows[2:]:
        rec = {"kernel": re.sub(r"\(.*", "", r[idx["Kernel Name"]]).replace("void ", "")}
        for m in METRICS:
            if m in idx:
                rec[m] = (r[idx[m]], units[idx[m]])
        out.append(rec)
    return out


def targets_summary(reps, labels):
    lines = ["`ncu --set full --clock-control none --import-source on --profile-from-start off` of tools/ncu_targets.py: ONE launch of every hot",
             "kernel at its cfg3 / cfg5 shape (cold caches: ncu flushes L2 between replays). Algorithmic bytes / FLOPs per launch in the last column.",
             ""]
    traffic = {}
    for rep, label in zip(reps, labels):
        if not rep.exists():
            continue
        lines.append(f"== {rep.name} ==")
        lines.append(f"{'kernel':44s} {'what':34s} {'us':>8s} {'dram rd MB':>10s} {'dram wr MB':>10s} {'dram %':>7s} {'tensor %':>8s} {'xu %':>6s} "
                     f"{'L2 hit %':>8s} {'regs':>5s} {'grid':>6s} {'GHz':>5s}  achieved")
        for rec, (what, alg) in zip(ncu_rows(rep), label):
            us = to_us(*rec["gpu__time_duration.sum"])
            rd, wr = to_bytes(*rec["dram__bytes_read.sum"]), to_bytes(*rec["dram__bytes_write.sum"])
            ach = ""
            if alg.get("bytes"):
                ach = f"{alg['bytes'] / us / 1e6:.2f} TB/s algorithmic ({alg['bytes'] / 1e6:.1f} MB)"
            if alg.get("flops"):
                ach = f"{alg['flops'] / us / 1e6:.0f} TFLOP/s ({alg['flops'] / 1e9:.1f} GFLOP)"
            g = lambda m: rec.get(m, ("", ""))[0]  # noqa: E731
            lines.append(f"{rec['kernel'][:44]:44s} {what[:34]:34s} {us:8.1f} {rd / 1e6:10.1f} {wr / 1e6:10.1f} "
                         f"{float(g('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed') or 0):7.1f} "
                         f"{float(g('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active') or 0):8.1f} "
                         f"{float(g('sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active') or 0):6.1f} "
                         f"{float(g('lts__t_sector_hit_rate.pct') or 0):8.1f} {g('launch__registers_per_thread'):>5s} {g('launch__grid_size'):>6s} "
                         f"{float(g('sm__cycles_elapsed.avg.per_second') or 0):5.2f}  {ach}")
            if "proj_mlp" in what:
                traffic["gemm"] = {"dram_bytes_read": rd, "dram_bytes_write": wr, "duration_us": us,
                                   "note": f"{rec['kernel']} {what}: algorithmic 28.3 MB (A) + 75.5 MB (W) + 113.2 MB (C) = 217 MB per launch; "
                                           "C stays in the 126 MB L2 as dirty lines within the kernel's duration"}
        lines.append("")
    return "\n".join(lines), traffic


def sass_histogram():
    so = ROOT / "unigen_b200" / "libunigen_b200.so"
    txt = subprocess.run(["cuobjdump", "-sass", str(so)], capture_output=True, text=True).stdout
    hist, cur = collections.OrderedDict(), None
    for ln in txt.splitlines():
        m = re.match(r"\s*Function : (\S+)", ln)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = re.sub(r"\(.*", "", name).replace("void ", "")
            hist[cur] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", ln)
        if m and cur:
            hist[cur][m.group(1)] += 1
    keys = ["UTCHMMA", "UTCQMMA", "UTMALDG", "UTMASTG", "LDTM", "STTM", "UTCBAR", "UTCCP", "SYNCS", "MUFU", "HMMA", "LDG", "STG", "LDS", "STS"]
    lines = [f"cuobjdump -sass {so.name}: opcode-family counts per kernel (UTCHMMA = tcgen05.mma, UTMALDG / UTMASTG = TMA load / store,",
             "LDTM / STTM = tcgen05.ld / st, UTCBAR = tcgen05.commit, SYNCS = mbarrier, HMMA = legacy mma.sync — must be 0)", "",
             f"{'kernel':70s} " + " ".join(f"{k:>8s}" for k in keys) + "    total"]
    for name, c in hist.items():
        fam = {k: sum(v for op, v in c.items() if op.split(".")[0].startswith(k)) for k in keys}
        lines.append(f"{name[:70]:70s} " + " ".join(f"{fam[k]:8d}" for k in keys) + f" {sum(c.values()):8d}")
    return "\n".join(lines) + "\n"


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r02a"
    PROF.mkdir(exist_ok=True)
    for f in OUT.glob(f"{tag}_bench_*.json"):
        shutil.copy(f, PROF / f.name.replace(tag, "r02"))
    launches = OUT / f"{tag}_launches_cfg3.csv"
    if launches.exists():
        shutil.copy(launches, PROF / "r02_launches_cfg3.csv")
        summ = subprocess.run([sys.executable, str(ROOT / "tools" / "launch_summary.py"), str(launches)], capture_output=True, text=True).stdout
        (PROF / "r02_launches_cfg3_summary.txt").write_text(
            "ncu --metrics gpu__time_duration.sum --clock-control none -k regex:ug:: of `python bench.py --steps 1 --warmup 1 --no-graph` "
            "(warm-up steps included; cold-cache, serialised launches: compare SHARES, not absolute times)\n" + summ)
    S, D, N = 4608, 3072, 4096
    cfg3 = [("grouped AdaLN GEMV (late table)", {"bytes": 9.157e9}), ("LN-modulate 4608 x 3072", {"bytes": 4.0 * S * D}),
            ("q|k|v GEMM 4608x3072->9216", {"flops": 2.0 * S * D * 3 * D}), ("q|k|v GEMM + fused QK-norm+RoPE", {"flops": 2.0 * S * D * 3 * D}),
            ("q|k|v GEMM (repeat, plain)", {"flops": 2.0 * S * D * 3 * D}), ("QK-RMSNorm+RoPE separate pass 4608x6144", {"bytes": 4.0 * S * 2 * D}),
            ("attention S=4608 H=24 dh=128", {"flops": 4.0 * S * S * D}), ("proj_mlp GEMM + GELU 4608x3072->12288", {"flops": 2.0 * S * D * 4 * D}),
            ("proj_out GEMM 4608x15360->3072 gate+res", {"flops": 2.0 * S * 5 * D * D}), ("ff2 GEMM 4096x12288->3072 gate+res", {"flops": 2.0 * N * 4 * D * D})]
    S5, D5 = 4096 + 333, 1536
    cfg5 = [("q|k|v GEMM 4429x1536->4608 (B=2)", {"flops": 2 * 2.0 * S5 * D5 * 3 * D5}), ("attention S=4429 H=24 dh=64 (B=2)", {"flops": 2 * 4.0 * S5 * S5 * D5})]
    text, traffic = targets_summary([OUT / "prof_r02_targets.ncu-rep", OUT / "prof_r02_targets_sd3.ncu-rep"], [cfg3, cfg5])
    if traffic:
        (PROF / "r02_ncu_targets_summary.txt").write_text(text)
        (PROF / "r02_ncu_traffic.json").write_text(json.dumps(traffic, indent=1) + "\n")
    (PROF / "r02_sass_opcodes.txt").write_text(sass_histogram())
    print("profiles/ updated:", sorted(p.name for p in PROF.glob("r02_*")))


if __name__ == "__main__":
    main()
