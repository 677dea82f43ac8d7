#!/bin/bash
# Round-2 GPU call B (1 GPU): re-run the GPU suite, then same-box A/B of the staged GEMM epilogue and the rows LayerNorm kernel.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/r02b_smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/r02b_smoke.log
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --durations=10 > gpurun_out/r02b_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02b_pytest.log
tail -12 gpurun_out/r02b_pytest.log
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline"
UG_GEMM_EPILOGUE=direct UG_LN_KERNEL=warp timeout 300 $B > gpurun_out/r02b_ab_direct_warp.json 2> gpurun_out/r02b_ab_direct_warp.err; echo "A exit $?"
UG_GEMM_EPILOGUE=direct UG_LN_KERNEL=rows timeout 300 $B > gpurun_out/r02b_ab_direct_rows.json 2> gpurun_out/r02b_ab_direct_rows.err; echo "B exit $?"
UG_GEMM_EPILOGUE=staged UG_LN_KERNEL=rows timeout 300 $B > gpurun_out/r02b_ab_staged_rows.json 2> gpurun_out/r02b_ab_staged_rows.err; echo "C exit $?"
UG_GEMM_EPILOGUE=direct UG_LN_KERNEL=warp timeout 300 $B > gpurun_out/r02b_ab_direct_warp2.json 2> gpurun_out/r02b_ab_direct_warp2.err; echo "A2 exit $?"
UG_GEMM_EPILOGUE=staged UG_LN_KERNEL=rows timeout 300 $B > gpurun_out/r02b_ab_staged_rows2.json 2> gpurun_out/r02b_ab_staged_rows2.err; echo "C2 exit $?"
for f in gpurun_out/r02b_ab_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d["roofline"]
    print(sys.argv[1].split("/")[-1], "ms/step %.2f"%d["ms_per_step"], "gemm %.1f TF %.2f ms"%(r["achieved"],r["ms_per_step_in_kernel"]), "attn %.2f ms"%r["attention"]["ms_per_step_in_kernel"],
          "ln %.2f ms"%r["hbm_bound_kernels"]["ln_modulate"]["ms_per_step_in_kernel"], "clk", d["clocks"]["sm_mhz"])
except Exception as e: print(sys.argv[1], "FAILED", e)
PY
done
