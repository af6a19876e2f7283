#!/bin/bash
# round 2, call 22: two-ring halo conv form (HALO=1) parity + timing; 64-wide tiles with alternating epilogue groups
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
run() { local name="$1" t="$2"; shift 2; echo "=== $name (timeout ${t}s)"; local t0=$SECONDS
    timeout -k 5 "$t" "$@" > "gpurun_out/$name.log" 2>&1; local rc=$?
    echo "=== $name rc=$rc $((SECONDS - t0))s"; tail -n "${TAILN:-12}" "gpurun_out/$name.log" | cut -c1-300; }
TAILN=14 run tests_conv 300 python -u -m pytest -q -m gpu -p no:cacheprovider --timeout 120 -rf tests/test_lpips_gpu.py
for h in 0 1; do
    export VTP_GEMM_CONV_HALO=$h
    TAILN=15 run lpips_layers_halo$h 90 python tools/lpips_layers_bench.py
done
for h in 0 1; do
    export VTP_GEMM_CONV_HALO=$h
    TAILN=1 run bench_halo$h 200 python bench.py --steps 8 --warmup 3 --no-cpu-baseline
    grep -E "device-resident|end-to-end" gpurun_out/bench_halo$h.log
done
