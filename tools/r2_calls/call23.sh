#!/bin/bash
# round 2, call 23: attention TMA-store epilogues (fwd pipe FULL, bwd FULL) + O-row prefetch: parity + timing
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
run() { local name="$1" t="$2"; shift 2; echo "=== $name (timeout ${t}s)"; local t0=$SECONDS
    timeout -k 5 "$t" "$@" > "gpurun_out/$name.log" 2>&1; local rc=$?
    echo "=== $name rc=$rc $((SECONDS - t0))s"; tail -n "${TAILN:-12}" "gpurun_out/$name.log" | cut -c1-400; }
TAILN=8 run attn_prof 120 python tools/attn_prof.py
export VTP_ATTN_PIPE_TSO=1 VTP_ATTN_BWD_TSO=1
TAILN=8 run tests_attn_tso 400 python -u -m pytest -q -m gpu -p no:cacheprovider --timeout 180 -rf tests/test_kernels_gpu.py tests/test_backward_gpu.py tests/test_model_gpu.py tests/test_train_gpu.py
TAILN=1 run bench_tso1 200 python bench.py --steps 8 --warmup 3 --no-cpu-baseline
grep -E "device-resident|end-to-end" gpurun_out/bench_tso1.log
export VTP_ATTN_PIPE_TSO=0 VTP_ATTN_BWD_TSO=0
TAILN=1 run bench_tso0 200 python bench.py --steps 8 --warmup 3 --no-cpu-baseline
grep -E "device-resident|end-to-end" gpurun_out/bench_tso0.log
