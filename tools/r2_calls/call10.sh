#!/bin/bash
# BASELINE configs[2]: VTP-Base full step, batch 1024 over 4 GPUs (256/GPU), contrastive feature all-gather across 4
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
run() { local name="$1" t="$2"; shift 2; echo "=== $name (timeout ${t}s)"; local t0=$SECONDS
    timeout -k 5 "$t" "$@" > "gpurun_out/$name.log" 2>&1; local rc=$?
    echo "=== $name rc=$rc $((SECONDS - t0))s"; tail -n "${TAILN:-6}" "gpurun_out/$name.log" | cut -c1-2500; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
run bench_base_n4 200 $TR --master-port 29521 bench.py --gpus 4 --model base --steps 6 --warmup 3
VTP_CLIP_EXCHANGE=p2p run bench_base_n4_p2p 200 $TR --master-port 29522 bench.py --gpus 4 --model base --steps 6 --warmup 3
