#!/bin/bash
# Round-2 GPU call A (1 GPU): full GPU test suite, parity records, bench lines, launch list, ncu captures.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export UG_PARITY_OUT=gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/r02a_gpu.txt 2>&1
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/r02a_smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/r02a_smoke.log
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --durations=15 > gpurun_out/r02a_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02a_pytest.log
tail -5 gpurun_out/r02a_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r02a_bench_cfg3.json 2> gpurun_out/r02a_bench_cfg3.err; echo "bench exit $?"
timeout 300 python bench.py --steps 3 --warmup 3 --batch 8 --no-cpu-baseline --no-eager-baseline > gpurun_out/r02a_bench_cfg3_b8.json 2> gpurun_out/r02a_bench_cfg3_b8.err; echo "bench b8 exit $?"
timeout 300 python bench.py --workload cfg2 --steps 8 --warmup 3 --loop 4 --no-cpu-baseline > gpurun_out/r02a_bench_cfg2_loop.json 2> gpurun_out/r02a_bench_cfg2_loop.err; echo "bench cfg2 loop exit $?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:ug:: -c 4000 --csv \
  --log-file gpurun_out/r02a_launches_cfg3.csv python bench.py --steps 1 --warmup 1 --no-graph --no-cpu-baseline --no-eager-baseline > gpurun_out/r02a_ncu_launches.log 2>&1; echo "ncu launches exit $?"
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/prof_r02_targets \
  python tools/ncu_targets.py > gpurun_out/r02a_ncu_targets.log 2>&1; echo "ncu targets exit $?"
timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/prof_r02_targets_sd3 \
  python tools/ncu_targets.py --workload cfg5 > gpurun_out/r02a_ncu_targets_sd3.log 2>&1; echo "ncu targets sd3 exit $?"
ls -la gpurun_out | tail -20
