#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
run() { local name="$1" t="$2"; shift 2; echo "=== $name (timeout ${t}s)"; local t0=$SECONDS
    timeout -k 5 "$t" "$@" > "gpurun_out/$name.log" 2>&1; local rc=$?
    echo "=== $name rc=$rc $((SECONDS - t0))s"; tail -n "${TAILN:-12}" "gpurun_out/$name.log" | cut -c1-700; }
TAILN=8 run eager_ctx 240 python tests/eager_gpu_context.py --batches 32,64,128 --steps 3
