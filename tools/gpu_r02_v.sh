#!/bin/bash
# ncu --set full capture of the VAE kernels (one launch each, tools/ncu_targets.py --workload vae) + a launch list of one decode
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python tools/ncu_targets.py --workload vae; echo "plain run exit $?"
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/prof_r02_targets_vae \
  python tools/ncu_targets.py --workload vae > gpurun_out/r02v_ncu_targets_vae.log 2>&1; echo "ncu targets vae exit $?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:ug:: -c 2000 --csv \
  --log-file gpurun_out/r02v_launches_vae.csv python tools/bench_vae.py --steps 1 --warmup 0 > gpurun_out/r02v_ncu_launches_vae.log 2>&1; echo "ncu launches exit $?"
ls -la gpurun_out | tail -5
