#!/bin/bash
# 8 GPUs: gradient all-reduce overlapped with backward vs after the last backward (and with NCCL limited to 8 CTAs)
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
run() { local name="$1" t="$2"; shift 2; echo "=== $name (timeout ${t}s)"; local t0=$SECONDS
    timeout -k 5 "$t" "$@" > "gpurun_out/$name.log" 2>&1; local rc=$?
    echo "=== $name rc=$rc $((SECONDS - t0))s"; grep -E "device-resident|end-to-end" "gpurun_out/$name.log"; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
VTP_GRAD_REDUCE=end run small_n8_end 150 $TR --master-port 29541 bench.py --gpus 8 --steps 10 --warmup 3
VTP_GRAD_REDUCE=overlap NCCL_MAX_CTAS=8 run small_n8_overlap_cta8 150 $TR --master-port 29542 bench.py --gpus 8 --steps 10 --warmup 3
VTP_GRAD_REDUCE=end run large_n8_end 250 $TR --master-port 29543 bench.py --gpus 8 --model large --steps 4 --warmup 2
