#!/usr/bin/env bash
set -u
N=${1:-2}
mkdir -p gpurun_out
P=29580
run() { P=$((P+1)); timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P "$@"; }
run tools/ddp_timeline.py --out gpurun_out/ddp_timeline_n${N} > gpurun_out/tl.log 2>&1; echo "timeline rc=$?"; head -n 4 gpurun_out/ddp_timeline_n${N}.txt | cut -c1-1500 || tail -n 5 gpurun_out/tl.log
NCCL_GRAPH_REGISTER=0 run tools/ddp_timeline.py --out gpurun_out/ddp_timeline_n${N}_noreg > gpurun_out/tl_noreg.log 2>&1; echo "timeline noreg rc=$?"; head -n 4 gpurun_out/ddp_timeline_n${N}_noreg.txt | cut -c1-1500
timeout 300 python tools/ddp_timeline.py --out gpurun_out/ddp_timeline_n1 > gpurun_out/tl_n1.log 2>&1; echo "timeline n1 rc=$?"; head -n 4 gpurun_out/ddp_timeline_n1.txt | cut -c1-1500
