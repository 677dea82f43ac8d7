#!/bin/bash
# round-end rehearsal on one GPU: the driver's three commands (pytest -m gpu, smoke, bench both arms), each timed
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
t0=$(date +%s)
timeout 2400 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider --durations=15 > gpurun_out/r02s_pytest.log 2>&1; echo "pytest exit $? after $(( $(date +%s) - t0 )) s" | tee -a gpurun_out/r02s_pytest.log
tail -30 gpurun_out/r02s_pytest.log | cut -c1-250
t0=$(date +%s)
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02s_smoke.log 2>&1; echo "smoke exit $? after $(( $(date +%s) - t0 )) s"
tail -3 gpurun_out/r02s_smoke.log | cut -c1-300
t0=$(date +%s)
timeout 900 python bench.py --impl reference --gpus 1 --steps 5 --warmup 3 > gpurun_out/r02s_bench_reference.json 2> gpurun_out/r02s_bench_reference.err; echo "reference arm exit $? after $(( $(date +%s) - t0 )) s"
tail -c 500 gpurun_out/r02s_bench_reference.json
t0=$(date +%s)
timeout 900 python bench.py --gpus 1 --steps 5 --warmup 3 > gpurun_out/r02s_bench_cfg3.json 2> gpurun_out/r02s_bench_cfg3.err; echo "native arm exit $? after $(( $(date +%s) - t0 )) s"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02s_bench_cfg3.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step', 'gpu_launches', 'clocks')}, d['e2e'], d['roofline']['frac'], d['roofline']['attention'])
PY
