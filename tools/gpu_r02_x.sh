#!/bin/bash
# 8 GPUs: the driver's N = 8 command on the final build (data-parallel value + the sequence-parallel legs)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L | wc -l
t0=$(date +%s)
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r02x_bench_cfg3_n8.json 2> gpurun_out/r02x_bench_cfg3_n8.err; echo "bench n8 exit $? after $(( $(date +%s) - t0 )) s"
tail -3 gpurun_out/r02x_bench_cfg3_n8.err | cut -c1-300
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02x_bench_cfg3_n8.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step', 'n_gpus')}, json.dumps(d.get('sp'))[:1500])
PY
