#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
run() { local name="$1" t="$2"; shift 2; echo "=== $name (timeout ${t}s)"; local t0=$SECONDS
    timeout "$t" "$@" > "gpurun_out/$name.log" 2>&1; local rc=$?
    echo "=== $name rc=$rc $((SECONDS - t0))s"; tail -n "${TAILN:-12}" "gpurun_out/$name.log"; }
TAILN=30 run tests_attn 300 python -u -m pytest -q -m gpu -p no:cacheprovider --timeout 120 -rfs tests/test_kernels_gpu.py tests/test_backward_gpu.py
run attn_prof 60 python tools/attn_prof.py
VTP_ATTN_PIPE_GENERIC=1 VTP_ATTN_BWD_GENERIC=1 run attn_prof_generic 60 python tools/attn_prof.py
run attn_prof_dec 60 python tools/attn_prof.py 256 256
TAILN=30 run tests_all 400 python -u -m pytest -q -m gpu -p no:cacheprovider --timeout 120 -x -rfs tests
TAILN=4 run bench_graph 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline
