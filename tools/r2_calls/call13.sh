#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
run() { local name="$1" t="$2"; shift 2; echo "=== $name (timeout ${t}s)"; local t0=$SECONDS
    timeout -k 5 "$t" "$@" > "gpurun_out/$name.log" 2>&1; local rc=$?
    echo "=== $name rc=$rc $((SECONDS - t0))s"; tail -n "${TAILN:-12}" "gpurun_out/$name.log" | cut -c1-600; }
TAILN=25 run tests_swiglu 300 python -u -m pytest -q -m gpu -p no:cacheprovider --timeout 120 -x -rfs tests/test_gemm_gpu.py tests/test_model_gpu.py tests/test_train_gpu.py tests/test_chunk_gpu.py tests/test_graphs_gpu.py
TAILN=20 run gemm_bench 90 python tools/gemm_bench.py
run bench_fused 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline
grep -E "device-resident|end-to-end" gpurun_out/bench_fused.log
VTP_FUSED_SWIGLU=0 run bench_split 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline
grep -E "device-resident|end-to-end" gpurun_out/bench_split.log
