#!/bin/bash
# 2-GPU call: distributed tests + N=2 bench (graph / eager / peer-memory exchange)
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
run() { local name="$1" t="$2"; shift 2; echo "=== $name (timeout ${t}s)"; local t0=$SECONDS
    timeout "$t" "$@" > "gpurun_out/$name.log" 2>&1; local rc=$?
    echo "=== $name rc=$rc $((SECONDS - t0))s"; tail -n "${TAILN:-12}" "gpurun_out/$name.log"; }
nvidia-smi -L
VTP_TEST_UNVALIDATED=1 TAILN=25 run tests_dist2 400 python -u -m pytest -q -m gpu -p no:cacheprovider --timeout 300 -rfs tests/test_dist_gpu.py
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
TAILN=3 run bench_n2_graph 300 $TR --master-port 29511 bench.py --gpus 2 --steps 8 --warmup 3
TAILN=3 run bench_n2_eager 300 $TR --master-port 29512 bench.py --gpus 2 --steps 8 --warmup 3 --graph off
VTP_CLIP_EXCHANGE=p2p TAILN=3 run bench_n2_p2p 300 $TR --master-port 29513 bench.py --gpus 2 --steps 8 --warmup 3
TAILN=30 run tests_train 400 python -u -m pytest -q -m gpu -p no:cacheprovider --timeout 180 -rfs -s tests/test_train_gpu.py
grep -E "^\[(tiny|small2|large2)\]|stochastic-depth rec" gpurun_out/tests_train.log | cut -c1-250
