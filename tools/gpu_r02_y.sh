#!/bin/bash
# final rehearsal: full GPU suite + smoke + cfg5 step (head_dim-64 attention default = variant 9) + cfg3 step
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/r02y_pytest.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/r02y_pytest.log
tail -4 gpurun_out/r02y_pytest.log | cut -c1-250
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 300 python tools/bench_sd3.py --batch 4 --steps 5 > gpurun_out/r02y_bench_sd3_b4.json 2> gpurun_out/r02y_bench_sd3.err; echo "bench sd3 exit $?"
timeout 300 python tools/bench_sd3.py --batch 4 --steps 5 --attn-variant 5 > gpurun_out/r02y_bench_sd3_b4_attn5.json 2>> gpurun_out/r02y_bench_sd3.err; echo "bench sd3 attn5 exit $?"
python - <<'PY'
import json
for f in ("r02y_bench_sd3_b4", "r02y_bench_sd3_b4_attn5"):
    d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
    print(f, d["ms_per_step"])
PY
