#!/bin/bash
# P-variant: fused QK-norm in the (K-extension) projection GEMMs — tests (1 GPU, or world-2 when 2 GPUs) + cfg4-P step
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_pvariant_gpu.py tests/test_parallel_gpu.py tests/test_fullsize_gpu.py -m gpu -q -p no:cacheprovider -k "not cfg3 and not cfg2 and not sd3" > gpurun_out/r02q_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02q_pytest.log
tail -5 gpurun_out/r02q_pytest.log | cut -c1-200
timeout 300 python tools/bench_pvariant.py --steps 4 > gpurun_out/r02q_bench_cfg4_pvariant.json 2> gpurun_out/r02q_bench_pv.err; echo "bench pv exit $?"
python -c "
import json; d=json.loads(open('gpurun_out/r02q_bench_cfg4_pvariant.json').read().strip().splitlines()[-1]); print({k:d[k] for k in d if k in ('ms_per_step','model_tflops','gpu_launches')})"
