#!/bin/bash
# VAE: full test file + bench at 1024^2 (with the torch-eager comparator) + 512^2
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_vae_gpu.py -m gpu -q -p no:cacheprovider > gpurun_out/r02u_pytest_vae.log 2>&1; echo "pytest vae exit $?" >> gpurun_out/r02u_pytest_vae.log
tail -25 gpurun_out/r02u_pytest_vae.log | cut -c1-400
timeout 600 python tools/bench_vae.py --side 1024 --eager > gpurun_out/r02u_bench_vae_1024.json 2> gpurun_out/r02u_bench_vae_1024.err; echo "bench vae exit $?"
tail -3 gpurun_out/r02u_bench_vae_1024.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02u_bench_vae_1024.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('decode_ms', 'encode_ms', 'decode_tflops', 'encode_tflops', 'decode_launches')}, d.get('eager_bf16'))
for k, v in d['decode_ops'].items():
    print('dec', k, v['launches'], round(v['ms'], 3), round(v['achieved'], 1), v['unit'])
for k, v in d['encode_ops'].items():
    print('enc', k, v['launches'], round(v['ms'], 3), round(v['achieved'], 1), v['unit'])
PY
