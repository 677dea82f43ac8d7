#!/bin/bash
# Round-2 evidence capture on the final build (1 GPU): full GPU suite, bench lines, launch list, ncu captures.
set -u
cd "$(dirname "$0")/.."
T=${1:-r02f}
mkdir -p gpurun_out
export UG_PARITY_OUT=gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/${T}_gpu.txt 2>&1
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/${T}_smoke.log
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --durations=10 > gpurun_out/${T}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${T}_pytest.log
tail -6 gpurun_out/${T}_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/${T}_bench_cfg3.json 2> gpurun_out/${T}_bench_cfg3.err; echo "bench exit $?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_bench_cfg3_reference_arm.json 2> gpurun_out/${T}_bench_ref.err; echo "bench ref exit $?"
timeout 300 python bench.py --steps 3 --warmup 3 --batch 8 --no-cpu-baseline --no-eager-baseline > gpurun_out/${T}_bench_cfg3_b8.json 2> gpurun_out/${T}_bench_cfg3_b8.err; echo "bench b8 exit $?"
timeout 300 python bench.py --workload cfg2 --steps 8 --warmup 3 --loop 4 --no-cpu-baseline > gpurun_out/${T}_bench_cfg2_loop.json 2> gpurun_out/${T}_bench_cfg2_loop.err; echo "bench cfg2 loop exit $?"
timeout 300 python bench.py --workload cfg4 --steps 5 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/${T}_bench_cfg4_svariant.json 2> gpurun_out/${T}_bench_cfg4.err; echo "bench cfg4 exit $?"
timeout 300 python tools/bench_pvariant.py --steps 4 > gpurun_out/${T}_bench_cfg4_pvariant.json 2> gpurun_out/${T}_bench_pv.err; echo "bench pv exit $?"
timeout 300 python tools/bench_sd3.py --batch 4 --steps 5 > gpurun_out/${T}_bench_sd3_b4.json 2> gpurun_out/${T}_bench_sd3.err; echo "bench sd3 exit $?"
timeout 300 python tools/bench_vae.py --side 1024 --eager > gpurun_out/${T}_bench_vae_1024.json 2> gpurun_out/${T}_bench_vae.err; echo "bench vae exit $?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:ug:: -c 4000 --csv \
  --log-file gpurun_out/${T}_launches_cfg3.csv python bench.py --steps 1 --warmup 1 --no-graph --no-cpu-baseline --no-eager-baseline > gpurun_out/${T}_ncu_launches.log 2>&1; echo "ncu launches exit $?"
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/prof_r02_targets \
  python tools/ncu_targets.py > gpurun_out/${T}_ncu_targets.log 2>&1; echo "ncu targets exit $?"
timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/prof_r02_targets_sd3 \
  python tools/ncu_targets.py --workload cfg5 > gpurun_out/${T}_ncu_targets_sd3.log 2>&1; echo "ncu targets sd3 exit $?"
