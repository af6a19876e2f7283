#!/bin/bash
# round 2, call 25: ncu --set full of the final attention kernels and of the halo-block conv1_2 kernel
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 100 ncu --set full --clock-control none --import-source on -k regex:attn -o gpurun_out/attn_r2c -f python tools/attn_prof.py --once > gpurun_out/ncu_attn_r2c.log 2>&1
echo "attn rc=$?"
timeout 60 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -c 4 -o gpurun_out/conv1_r2c -f python tools/lpips_conv1_once.py > gpurun_out/ncu_conv1_r2c.log 2>&1
echo "conv rc=$?"
ls -la gpurun_out/*r2c*
