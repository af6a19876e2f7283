#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
run() { local name="$1" t="$2"; shift 2; echo "=== $name (timeout ${t}s)"; local t0=$SECONDS
    timeout -k 5 "$t" "$@" > "gpurun_out/$name.log" 2>&1; local rc=$?
    echo "=== $name rc=$rc $((SECONDS - t0))s"; tail -n "${TAILN:-12}" "gpurun_out/$name.log" | cut -c1-1200; }
TAILN=40 run tests_all 600 python -u -m pytest -q -m gpu -p no:cacheprovider --timeout 180 -rfs -s tests
grep -E "^\[(tiny|small2|large2)\]|stochastic" gpurun_out/tests_all.log | cut -c1-250
TAILN=3 run smoke 100 python -c "import __graft_entry__ as g; g.smoke()"
