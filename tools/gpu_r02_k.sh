#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:attention2 -c 1 -f -o gpurun_out/prof_r02k_attn python tools/ncu_targets.py > gpurun_out/r02k_ncu.log 2>&1; echo "ncu exit $?"
tail -3 gpurun_out/r02k_ncu.log
