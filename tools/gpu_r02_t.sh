#!/bin/bash
# VAE ops + model tests; GEMM op tests (the producer warp gained the implicit-convolution branch)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_vae_gpu.py -m gpu -q -p no:cacheprovider -x > gpurun_out/r02t_pytest_vae.log 2>&1; echo "pytest vae exit $?" >> gpurun_out/r02t_pytest_vae.log
tail -40 gpurun_out/r02t_pytest_vae.log | cut -c1-400
timeout 900 python -m pytest tests/test_ops_gpu.py -m gpu -q -p no:cacheprovider > gpurun_out/r02t_pytest_ops.log 2>&1; echo "pytest ops exit $?" >> gpurun_out/r02t_pytest_ops.log
tail -5 gpurun_out/r02t_pytest_ops.log | cut -c1-300
