#!/bin/bash
# round 2, call 21: resident-weight / halo-block conv forms (parity + timing), swiglu/gelu backward fix
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
run() { local name="$1" t="$2"; shift 2; echo "=== $name (timeout ${t}s)"; local t0=$SECONDS
    timeout -k 5 "$t" "$@" > "gpurun_out/$name.log" 2>&1; local rc=$?
    echo "=== $name rc=$rc $((SECONDS - t0))s"; tail -n "${TAILN:-12}" "gpurun_out/$name.log" | cut -c1-300; }
TAILN=14 run tests_conv 240 python -u -m pytest -q -m gpu -p no:cacheprovider --timeout 120 -rf tests/test_lpips_gpu.py -k conv_mode
export VTP_GEMM_HALO_BO=0
TAILN=8 run tests_conv_bo0 120 python -u -m pytest -q -m gpu -p no:cacheprovider --timeout 120 -rf tests/test_lpips_gpu.py -k "conv_mode_variants and BRES"
unset VTP_GEMM_HALO_BO
TAILN=6 run tests_kernels 240 python -u -m pytest -q -m gpu -p no:cacheprovider --timeout 120 -rf tests/test_kernels_gpu.py tests/test_backward_gpu.py
TAILN=16 run hbm_kernels 120 python tools/hbm_kernels_bench.py --out gpurun_out/hbm_kernels_r2e.json
for b in 0 1 2; do
    export VTP_GEMM_CONV_BRES=$b
    TAILN=4 run lpips_layers_bres$b 90 python tools/lpips_layers_bench.py
    grep -E "conv 1 " gpurun_out/lpips_layers_bres$b.log
done
for b in 0 1 2; do
    export VTP_GEMM_CONV_BRES=$b
    TAILN=1 run bench_bres$b 200 python bench.py --steps 8 --warmup 3 --no-cpu-baseline
    grep -E "device-resident|end-to-end" gpurun_out/bench_bres$b.log
done
