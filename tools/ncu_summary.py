"""Summarise an .ncu-rep: headline raw metrics + the hottest SASS instructions with their stall reasons.
Usage: python tools/ncu_summary.py <report.ncu-rep> [kernel-index] [min-samples]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
kidx = int(sys.argv[2]) if len(sys.argv) > 2 else 0
min_s = int(sys.argv[3]) if len(sys.argv) > 3 else 80
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
keys = ["Kernel Name", "gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_xu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__issue_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size", "launch__registers_per_thread",
        "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_tmem", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]
for i, h in enumerate(hdr):
    if any(h == k or h.endswith("." + k) or k in h and k.startswith("sm__inst_executed_pipe_tmem") for k in keys):
        print(f"{h[-86:]:86s} {units[i]:10s} " + " | ".join(r[i][:22] for r in rows[2:]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
secs, sec, h2, name = [], [], None, None
for r in rows:
    if r and r[0] == "Kernel Name":
        if sec:
            secs.append((name, h2, sec))
        name, sec, h2 = r[1], [], None
        continue
    if r and r[0] == "Address":
        h2 = r
        continue
    if h2:
        sec.append(r)
secs.append((name, h2, sec))
name, h2, sec = secs[kidx]
print("\n==", name[:100])
isrc, isamp = h2.index("Source"), h2.index("# Samples")
stall_cols = [i for i, h in enumerate(h2) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[isamp] or 0) for r in sec)
agg = {}
for r in sec:
    for i in stall_cols:
        agg[h2[i]] = agg.get(h2[i], 0) + int(r[i] or 0)
print("samples", tot, sorted(agg.items(), key=lambda x: -x[1])[:8])
for i, r in enumerate(sec):
    n = int(r[isamp] or 0)
    if n >= min_s:
        st = {h2[k]: int(r[k] or 0) for k in stall_cols if int(r[k] or 0) > 0}
        st = dict(sorted(st.items(), key=lambda x: -x[1])[:3])
        print(str(i).rjust(5), str(n).rjust(6), r[isrc][:64].ljust(64), st)
