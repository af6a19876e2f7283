#!/bin/bash
# 2-GPU verification of the graph + NCCL shutdown fix (short, strict timeouts)
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
run() { local name="$1" t="$2"; shift 2; echo "=== $name (timeout ${t}s)"; local t0=$SECONDS
    timeout -k 5 "$t" "$@" > "gpurun_out/$name.log" 2>&1; local rc=$?
    echo "=== $name rc=$rc $((SECONDS - t0))s"; tail -n "${TAILN:-12}" "gpurun_out/$name.log" | cut -c1-1500; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
TAILN=3 run bench_n2_graph 150 $TR --master-port 29511 bench.py --gpus 2 --steps 8 --warmup 3
VTP_CLIP_EXCHANGE=p2p TAILN=3 run bench_n2_p2p 150 $TR --master-port 29513 bench.py --gpus 2 --steps 8 --warmup 3
TAILN=12 run tests_dist2_graph 200 python -u -m pytest -q -m gpu -p no:cacheprovider --timeout 150 -rfs tests/test_dist_gpu.py -k graph
