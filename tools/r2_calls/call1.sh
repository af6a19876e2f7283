#!/bin/bash
# Round-2 first hardware pass (1 GPU): every gated case, the never-run model sizes, the opt-in kernel variants.
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
run() { local name="$1" t="$2"; shift 2; echo "=== $name (timeout ${t}s)"; local t0=$SECONDS
    timeout "$t" "$@" > "gpurun_out/$name.log" 2>&1; local rc=$?
    echo "=== $name rc=$rc $((SECONDS - t0))s"; tail -n "${TAILN:-12}" "gpurun_out/$name.log"; }
nvidia-smi --query-gpu=name,memory.total --format=csv
VTP_TEST_UNVALIDATED=1 TAILN=40 run tests_all 400 python -u -m pytest -q -m gpu -p no:cacheprovider --timeout 120 -k "not pipe" -rfs tests
run bench_large_n1 420 python bench.py --model large --steps 3 --warmup 2 --no-cpu-baseline
run bench_base_n1 240 python bench.py --model base --steps 3 --warmup 2 --no-cpu-baseline
TAILN=14 run infer_sweep_large 300 python tools/infer_sweep.py --model large --batches 1,2,4,8,16,32,64,128,256,512
for clm in 2 4 8; do VTP_GEMM_CLM=$clm TAILN=20 run gemm_bench_clm$clm 90 python tools/gemm_bench.py; done
VTP_TEST_UNVALIDATED=1 run attn_prof 60 python tools/attn_prof.py
VTP_TEST_UNVALIDATED=1 TAILN=25 run tests_pipe 120 python -u -m pytest -q -m gpu -p no:cacheprovider --timeout 60 -k "pipe" -rfs tests/test_kernels_gpu.py
