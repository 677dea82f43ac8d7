#!/bin/bash
# 2 GPUs: world-2 GPU tests of the sequence-parallel paths + the driver's N = 2 command (data-parallel value + SP legs)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L | head -4
timeout 900 python -m pytest tests/test_parallel_gpu.py -m gpu -q -p no:cacheprovider > gpurun_out/r02w_pytest_parallel.log 2>&1; echo "pytest parallel exit $?" >> gpurun_out/r02w_pytest_parallel.log
tail -6 gpurun_out/r02w_pytest_parallel.log | cut -c1-300
t0=$(date +%s)
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02w_bench_cfg3_n2.json 2> gpurun_out/r02w_bench_cfg3_n2.err; echo "bench n2 exit $? after $(( $(date +%s) - t0 )) s"
tail -3 gpurun_out/r02w_bench_cfg3_n2.err | cut -c1-300
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02w_bench_cfg3_n2.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step', 'n_gpus')}, json.dumps(d.get('sp'))[:1500])
PY
