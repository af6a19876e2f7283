#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
run() { local name="$1" t="$2"; shift 2; echo "=== $name (timeout ${t}s)"; local t0=$SECONDS
    timeout "$t" "$@" > "gpurun_out/$name.log" 2>&1; local rc=$?
    echo "=== $name rc=$rc $((SECONDS - t0))s"; tail -n "${TAILN:-12}" "gpurun_out/$name.log"; }
TAILN=40 run tests_all 600 python -u -m pytest -q -m gpu -p no:cacheprovider --timeout 180 -rfs -s tests
grep -E "grad rel errors: max" gpurun_out/tests_all.log | cut -c1-120
TAILN=30 run hbm_kernels 120 python tools/hbm_kernels_bench.py --out gpurun_out/hbm_kernels_r2b.json
TAILN=4 run bench_graph 300 python bench.py --steps 10 --warmup 3
TAILN=60 run step_gaps 200 python tools/step_gaps.py --graph --out gpurun_out/step_gaps_small_r2b.json
