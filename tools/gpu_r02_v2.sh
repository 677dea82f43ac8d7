#!/bin/bash
# ncu --set full capture of the VAE kernels on the final build (one launch each, tools/ncu_targets.py --workload vae)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 200 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/prof_r02_targets_vae \
  python tools/ncu_targets.py --workload vae > gpurun_out/r02v2_ncu_targets_vae.log 2>&1; echo "ncu targets vae exit $?"
