#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
run() { local name="$1" t="$2"; shift 2; echo "=== $name (timeout ${t}s)"; local t0=$SECONDS
    timeout "$t" "$@" > "gpurun_out/$name.log" 2>&1; local rc=$?
    echo "=== $name rc=$rc $((SECONDS - t0))s"; tail -n "${TAILN:-12}" "gpurun_out/$name.log"; }
TAILN=30 run tests_graph 300 python -u -m pytest -q -m gpu -p no:cacheprovider --timeout 120 -rfs tests/test_train_gpu.py tests/test_chunk_gpu.py tests/test_clip_gpu.py
TAILN=6 run bench_graph 300 python bench.py --steps 10 --warmup 3
TAILN=6 run bench_eager 300 python bench.py --steps 10 --warmup 3 --graph off --no-cpu-baseline
TAILN=6 run bench_large_graph 300 python bench.py --model large --steps 3 --warmup 2 --no-cpu-baseline
