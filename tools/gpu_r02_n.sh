#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_ops_gpu.py -m gpu -q -p no:cacheprovider -x -k "attention" > gpurun_out/r02n_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02n_pytest.log
tail -12 gpurun_out/r02n_pytest.log | cut -c1-220
UG_PROBE_VARIANTS=${VARIANTS:-7,5} timeout 600 python tools/probe_attn.py > gpurun_out/r02n_probe.log 2>&1; echo "probe exit $?"
grep -c '"ok": true' gpurun_out/r02n_probe.log; grep '"ok": false\|error\|TIMEOUT\|exit' gpurun_out/r02n_probe.log | cut -c1-400 | head
grep timing gpurun_out/r02n_probe.log | cut -c1-125
