#!/usr/bin/env bash
# N=2: bucket size and NCCL protocol against the step time (tools/ddp_timeline.py, 10 replays by CUDA events)
set -u
mkdir -p gpurun_out
P=29720
run() { P=$((P+1)); timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $P "$@"; }
for MB in 32 64 128 16; do
  run tools/ddp_timeline.py --bucket-mb $MB --out gpurun_out/ddp_tl_mb$MB > gpurun_out/tl_mb$MB.log 2>&1; echo "bucket $MB rc=$?"; head -n 1 gpurun_out/ddp_tl_mb$MB.txt | cut -c1-170; sed -n 3p gpurun_out/ddp_tl_mb$MB.txt
done
NCCL_PROTO=Simple run tools/ddp_timeline.py --out gpurun_out/ddp_tl_simple > gpurun_out/tl_simple.log 2>&1; echo "proto simple rc=$?"; head -n 1 gpurun_out/ddp_tl_simple.txt | cut -c1-170; sed -n 3p gpurun_out/ddp_tl_simple.txt
