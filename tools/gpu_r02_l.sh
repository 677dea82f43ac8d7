#!/bin/bash
# full GPU suite on the elect.sync / packed-softmax build + in-step A/B of the attention variants
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider -x > gpurun_out/r02l_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02l_pytest.log
tail -5 gpurun_out/r02l_pytest.log
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline"
for arm in 3 5 3 5; do
  timeout 300 $B --attn-variant $arm > gpurun_out/r02l_ab_attn${arm}_$RANDOM.json 2>> gpurun_out/r02l_ab.err; echo "arm $arm exit $?"
done
for f in gpurun_out/r02l_ab_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d["roofline"]
    print(sys.argv[1].split("/")[-1], "ms/step %.2f"%d["ms_per_step"], "gemm %.1f TF %.2f ms"%(r["achieved"], r["ms_per_step_in_kernel"]), "attn %.1f TF %.2f ms"%(r["attention"]["achieved"], r["attention"]["ms_per_step_in_kernel"]), "clk", d["clocks"]["sm_mhz"])
except Exception as e: print(sys.argv[1], "FAILED", e)
PY
done
tail -5 gpurun_out/r02l_ab.err
