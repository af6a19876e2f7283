#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 200 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -o gpurun_out/conv1_r2 -f python tools/lpips_conv1_once.py > gpurun_out/ncu_conv1.log 2>&1
tail -3 gpurun_out/ncu_conv1.log
for v in VTP_GEMM_CONV_NO_CLUSTER VTP_GEMM_CONV_NO_FAST; do echo "== $v"; env $v=1 timeout 60 python - <<'P'
import os, sys, torch
sys.path.insert(0, os.getcwd())
from vtp_b200 import lib
BF=torch.bfloat16
B,hw,ci,co=32,256,64,64
M=B*hw*hw
x=(torch.randn(B,hw,hw,ci,device="cuda")*0.5).to(BF); w=(torch.randn(co,9*ci,device="cuda")*0.05).to(BF); bias=torch.zeros(co,device="cuda"); y=torch.empty(B,hw,hw,co,device="cuda",dtype=BF)
f=lambda: lib.gemm(x,w,y,M=M,N=co,K=9*ci,lda=ci,ldb=9*ci,bias=bias,act=lib.ACT_RELU,ldo=co,conv=(ci,hw,hw))
for _ in range(3): f()
e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record()
for _ in range(10): f()
e1.record(); torch.cuda.synchronize()
print("conv1 fwd us:", e0.elapsed_time(e1)/10*1e3)
P
done
