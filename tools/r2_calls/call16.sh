#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
run() { local name="$1" t="$2"; shift 2; echo "=== $name (timeout ${t}s)"; local t0=$SECONDS
    timeout -k 5 "$t" "$@" > "gpurun_out/$name.log" 2>&1; local rc=$?
    echo "=== $name rc=$rc $((SECONDS - t0))s"; tail -n "${TAILN:-12}" "gpurun_out/$name.log" | cut -c1-300; }
TAILN=6 run tests_bw 300 python -u -m pytest -q -m gpu -p no:cacheprovider --timeout 120 -x -rfs tests/test_backward_gpu.py tests/test_train_gpu.py tests/test_kernels_gpu.py
TAILN=16 run hbm_kernels 120 python tools/hbm_kernels_bench.py --out gpurun_out/hbm_kernels_r2c.json
TAILN=16 run lpips_layers 120 python tools/lpips_layers_bench.py
TAILN=3 run bench 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline
