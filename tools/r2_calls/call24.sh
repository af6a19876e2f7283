#!/bin/bash
# round 2, call 24: full validation of the final build (halo convs, TMA-store attention epilogues, N = 64 resident-weight GEMM)
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
run() { local name="$1" t="$2"; shift 2; echo "=== $name (timeout ${t}s)"; local t0=$SECONDS
    timeout -k 5 "$t" "$@" > "gpurun_out/$name.log" 2>&1; local rc=$?
    echo "=== $name rc=$rc $((SECONDS - t0))s"; tail -n "${TAILN:-12}" "gpurun_out/$name.log" | cut -c1-300; }
TAILN=8 run tests_all 600 python -u -m pytest -q -m gpu -p no:cacheprovider --timeout 180 -rfs tests
TAILN=3 run smoke 100 python -c "import __graft_entry__ as g; g.smoke()"
TAILN=16 run hbm_kernels 120 python tools/hbm_kernels_bench.py --out gpurun_out/hbm_kernels_r2f.json
TAILN=5 run attn_prof 120 python tools/attn_prof.py
TAILN=16 run lpips_layers_final 90 python tools/lpips_layers_bench.py
TAILN=3 run bench_final 300 python bench.py --steps 20 --warmup 5
grep -E "device-resident|end-to-end" gpurun_out/bench_final.log
TAILN=50 run step_gaps 200 python tools/step_gaps.py --graph --out gpurun_out/step_gaps_small_r2d.json
TAILN=3 run ncu_launches 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 5200 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 1 --warmup 1 --graph off --no-cpu-baseline
ls -la gpurun_out/launches_r2.csv
