#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
run() { local name="$1" t="$2"; shift 2; echo "=== $name (timeout ${t}s)"; local t0=$SECONDS
    timeout "$t" "$@" > "gpurun_out/$name.log" 2>&1; local rc=$?
    echo "=== $name rc=$rc $((SECONDS - t0))s"; tail -n "${TAILN:-12}" "gpurun_out/$name.log"; }
TAILN=70 run step_gaps 200 python tools/step_gaps.py --out gpurun_out/step_gaps_small.json
VTP_TEST_UNVALIDATED=1 run ncu_attn 400 ncu --set full --clock-control none --import-source on -k regex:attn -o gpurun_out/attn_r2a -f python tools/attn_prof.py --once
ls -la gpurun_out/*.ncu-rep
