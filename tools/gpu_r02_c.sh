#!/bin/bash
# Round-2 GPU call C (N GPUs, default 2): sequence-parallel correctness (world-N tests, full-size bit-identity) + the driver-shaped bench line.
set -u
cd "$(dirname "$0")/.."
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02c_gpus_n$N.txt 2>&1
if [ "$N" = "2" ]; then
  timeout 900 python -m pytest tests/test_parallel_gpu.py tests/test_model_gpu.py::test_cuda_graph_replay_survives_alternating_shapes -m gpu -q -p no:cacheprovider > gpurun_out/r02c_pytest_n$N.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02c_pytest_n$N.log
  tail -5 gpurun_out/r02c_pytest_n$N.log
fi
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29611 tools/sp_check.py --workload cfg3 --steps 5 --exchange peer --graph > gpurun_out/r02c_sp_check_cfg3_n$N.json 2> gpurun_out/r02c_sp_check_cfg3_n$N.err; echo "sp_check cfg3 exit $?"
tail -c 1500 gpurun_out/r02c_sp_check_cfg3_n$N.json
timeout 600 $TR --master-port 29612 tools/sp_check_pvariant.py --workload cfg4 --steps 5 --graph --skip-single > gpurun_out/r02c_sp_check_cfg4p_n$N.json 2> gpurun_out/r02c_sp_check_cfg4p_n$N.err; echo "sp_check cfg4p exit $?"
tail -c 1200 gpurun_out/r02c_sp_check_cfg4p_n$N.json
timeout 900 $TR --master-port 29613 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02c_bench_n$N.json 2> gpurun_out/r02c_bench_n$N.err; echo "bench exit $?"
tail -c 3000 gpurun_out/r02c_bench_n$N.json
tail -5 gpurun_out/r02c_bench_n$N.err
