#!/bin/bash
# SD3 variants (use_modulate / use_shared_expert=False / init_control_param), UniGenSD3Pipeline entry: tests + cfg5 step
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sd3_gpu.py tests/test_pipeline_gpu.py -m gpu -q -p no:cacheprovider > gpurun_out/r02r_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02r_pytest.log
tail -25 gpurun_out/r02r_pytest.log | cut -c1-300
timeout 300 python bench.py --workload cfg5 --batch 4 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02r_bench_cfg5.json 2> gpurun_out/r02r_bench_cfg5.err; echo "bench cfg5 exit $?"
tail -c 600 gpurun_out/r02r_bench_cfg5.json
