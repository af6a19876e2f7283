#!/bin/bash
# BASELINE configs[3]: VTP-Large full step bf16, batch 2048 over 8 GPUs (256/GPU), gradient all-reduce over NVLink; + Small at N=8
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
run() { local name="$1" t="$2"; shift 2; echo "=== $name (timeout ${t}s)"; local t0=$SECONDS
    timeout -k 5 "$t" "$@" > "gpurun_out/$name.log" 2>&1; local rc=$?
    echo "=== $name rc=$rc $((SECONDS - t0))s"; tail -n "${TAILN:-6}" "gpurun_out/$name.log" | cut -c1-2500; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
run bench_large_n8 300 $TR --master-port 29531 bench.py --gpus 8 --model large --steps 4 --warmup 2
run bench_small_n8 200 $TR --master-port 29532 bench.py --gpus 8 --steps 10 --warmup 3
