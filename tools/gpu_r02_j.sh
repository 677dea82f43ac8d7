#!/bin/bash
# elect.sync MMA issue (attention v1/v2, GEMM) + packed softmax: correctness, isolated timing, in-step numbers
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
UG_PROBE_VARIANTS=${VARIANTS:-3,5,7} timeout 900 python tools/probe_attn.py > gpurun_out/r02j_probe.log 2>&1; echo "probe exit $?"
grep -c '"ok": true' gpurun_out/r02j_probe.log; grep '"ok": false\|error\|TIMEOUT' gpurun_out/r02j_probe.log | head
grep timing gpurun_out/r02j_probe.log | cut -c1-150
timeout 900 python -m pytest tests/test_ops_gpu.py -m gpu -q -p no:cacheprovider -x -k attention > gpurun_out/r02j_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02j_pytest.log
tail -4 gpurun_out/r02j_pytest.log
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline"
for arm in 0 7 0 7; do
  timeout 300 $B --attn-variant $arm > gpurun_out/r02j_ab_attn${arm}_$RANDOM.json 2>> gpurun_out/r02j_ab.err; echo "arm $arm exit $?"
done
for f in gpurun_out/r02j_ab_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d["roofline"]
    print(sys.argv[1].split("/")[-1], "ms/step %.2f"%d["ms_per_step"], "gemm %.1f TF %.2f ms"%(r["achieved"], r["ms_per_step_in_kernel"]), "attn %.1f TF %.2f ms"%(r["attention"]["achieved"], r["attention"]["ms_per_step_in_kernel"]), "clk", d["clocks"]["sm_mhz"])
except Exception as e: print(sys.argv[1], "FAILED", e)
PY
done
tail -5 gpurun_out/r02j_ab.err
