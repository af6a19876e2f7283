#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
run() { local name="$1" t="$2"; shift 2; echo "=== $name (timeout ${t}s)"; local t0=$SECONDS
    timeout -k 5 "$t" "$@" > "gpurun_out/$name.log" 2>&1; local rc=$?
    echo "=== $name rc=$rc $((SECONDS - t0))s"; tail -n "${TAILN:-12}" "gpurun_out/$name.log" | cut -c1-300; }
TAILN=5 run tests_gemm 300 python -u -m pytest -q -m gpu -p no:cacheprovider --timeout 120 -x -rfs tests/test_gemm_gpu.py tests/test_lpips_gpu.py tests/test_model_gpu.py tests/test_train_gpu.py
TAILN=18 run gemm_bench 90 python tools/gemm_bench.py
TAILN=3 run bench 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline
grep -E "device-resident|end-to-end" gpurun_out/bench.log
