#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
run() { local name="$1" t="$2"; shift 2; echo "=== $name (timeout ${t}s)"; local t0=$SECONDS
    timeout "$t" "$@" > "gpurun_out/$name.log" 2>&1; local rc=$?
    echo "=== $name rc=$rc $((SECONDS - t0))s"; tail -n "${TAILN:-12}" "gpurun_out/$name.log"; }
TAILN=5 run ncu_attn 300 ncu --set full --clock-control none --import-source on -k regex:attn -o gpurun_out/attn_r2b -f python tools/attn_prof.py --once
TAILN=5 run ncu_hbm 400 ncu --set full --clock-control none --import-source on -k 'regex:norm_|swiglu|rope|cast_colsum|patchify|adamw|dino' -o gpurun_out/hbm_r2 -f python tools/hbm_kernels_bench.py --once
TAILN=30 run hbm_kernels 120 python tools/hbm_kernels_bench.py --out gpurun_out/hbm_kernels_r2.json
for c in 64 128; do VTP_LPIPS_CHUNK=$c TAILN=3 run bench_lpips$c 200 python bench.py --steps 6 --warmup 2 --no-cpu-baseline; done
TAILN=8 run tests_gen 100 python -m pytest -q -m gpu -p no:cacheprovider tests/test_generation_gpu.py tests/test_graphs_gpu.py tests/test_model_gpu.py tests/test_lpips_gpu.py tests/test_gemm_gpu.py
ls -la gpurun_out/*.ncu-rep
