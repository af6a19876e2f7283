#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
run() { local name="$1" t="$2"; shift 2; echo "=== $name (timeout ${t}s)"; local t0=$SECONDS
    timeout -k 5 "$t" "$@" > "gpurun_out/$name.log" 2>&1; local rc=$?
    echo "=== $name rc=$rc $((SECONDS - t0))s"; tail -n "${TAILN:-12}" "gpurun_out/$name.log" | cut -c1-300; }
TAILN=8 run tests_all 600 python -u -m pytest -q -m gpu -p no:cacheprovider --timeout 180 -rfs tests
TAILN=16 run hbm_kernels 120 python tools/hbm_kernels_bench.py --out gpurun_out/hbm_kernels_r2d.json
TAILN=3 run smoke 100 python -c "import __graft_entry__ as g; g.smoke()"
TAILN=3 run bench_final 300 python bench.py --steps 20 --warmup 5
grep -E "device-resident|end-to-end" gpurun_out/bench_final.log
TAILN=60 run step_gaps 200 python tools/step_gaps.py --graph --out gpurun_out/step_gaps_small_r2c.json
