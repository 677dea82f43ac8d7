#!/bin/bash
# LayerNorm kernel A/B inside the cfg3 step (UG_LN_KERNEL=stream|rows) + the new tests
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_model_gpu.py -m gpu -q -p no:cacheprovider -k "ln_modulate or consis or tiny_forward" > gpurun_out/r02o_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02o_pytest.log
tail -3 gpurun_out/r02o_pytest.log
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline"
for arm in rows stream rows stream; do
  UG_LN_KERNEL=$arm timeout 300 $B > gpurun_out/r02o_ab_ln_${arm}_$RANDOM.json 2>> gpurun_out/r02o_ab.err; echo "arm $arm exit $?"
done
for f in gpurun_out/r02o_ab_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d["roofline"]; ln=r["hbm_bound_kernels"]["ln_modulate"]
    print(sys.argv[1].split("/")[-1], "ms/step %.2f"%d["ms_per_step"], "gemm %.1f TF"%r["achieved"], "attn %.1f TF"%r["attention"]["achieved"], "ln %.2f ms %.0f GB/s"%(ln["ms_per_step_in_kernel"], ln["achieved_gbs"]), "clk", d["clocks"]["sm_mhz"])
except Exception as e: print(sys.argv[1], "FAILED", e)
PY
done
tail -3 gpurun_out/r02o_ab.err
