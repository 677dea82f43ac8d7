#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_model_gpu.py tests/test_sd3_gpu.py -m gpu -q -p no:cacheprovider -k "gemm or fused or tiny_forward or ragged" > gpurun_out/r02e_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02e_pytest.log
tail -6 gpurun_out/r02e_pytest.log
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline"
for arm in 0 1 0 1; do
  timeout 300 $B --fuse-qk-norm $arm > gpurun_out/r02e_ab_fuseqk${arm}_$RANDOM.json 2>> gpurun_out/r02e_ab.err; echo "arm $arm exit $?"
done
for f in gpurun_out/r02e_ab_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d["roofline"]
    print(sys.argv[1].split("/")[-1], "ms/step %.2f"%d["ms_per_step"], "gemm %.1f TF %.2f ms"%(r["achieved"],r["ms_per_step_in_kernel"]), "attn %.2f ms"%r["attention"]["ms_per_step_in_kernel"],
          "qk %.2f ms (%d launches)"%(r["hbm_bound_kernels"]["qk_rmsnorm_rope"]["ms_per_step_in_kernel"], r["hbm_bound_kernels"]["qk_rmsnorm_rope"]["launches_per_step"]), "clk", d["clocks"]["sm_mhz"], "launches", d["gpu_launches"])
except Exception as e: print(sys.argv[1], "FAILED", e)
PY
done
tail -5 gpurun_out/r02e_ab.err
