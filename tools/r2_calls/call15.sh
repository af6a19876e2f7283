#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
run() { local name="$1" t="$2"; shift 2; echo "=== $name (timeout ${t}s)"; local t0=$SECONDS
    timeout -k 5 "$t" "$@" > "gpurun_out/$name.log" 2>&1; local rc=$?
    echo "=== $name rc=$rc $((SECONDS - t0))s"; tail -n "${TAILN:-12}" "gpurun_out/$name.log" | cut -c1-700; }
TAILN=8 run eager_ctx 240 python tests/eager_gpu_context.py --batches 32,64,128 --steps 3
TAILN=14 run tests_all 600 python -u -m pytest -q -m gpu -p no:cacheprovider --timeout 180 -rfs -s tests
grep -E "bf16-mode rel|ours vs fp32" gpurun_out/tests_all.log | cut -c1-260
TAILN=4 run ncu_gemm 200 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -o gpurun_out/gemm_r2 -f python tools/gemm_fc1_once.py
TAILN=3 run bench_final 300 python bench.py --steps 20 --warmup 5
TAILN=3 run bench_ref 300 python bench.py --impl reference --steps 20 --warmup 5
