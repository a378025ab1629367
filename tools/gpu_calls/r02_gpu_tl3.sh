#!/usr/bin/env bash
set -u
N=${1:-2}
mkdir -p gpurun_out
P=29620
run() { P=$((P+1)); timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P "$@"; }
for T in 1; do
PVQA_TAIL_PG=$T run tools/ddp_timeline.py --out gpurun_out/ddp_timeline_n${N}_tailpg$T > gpurun_out/tl_t$T.log 2>&1; echo "timeline tailpg $T rc=$?"; head -n 1 gpurun_out/ddp_timeline_n${N}_tailpg$T.txt | cut -c1-200 || tail -n 5 gpurun_out/tl_t$T.log
grep -c "AccumulateGrad node's stream" gpurun_out/tl_t$T.log
python - <<PY
import json
b = json.load(open("gpurun_out/ddp_timeline_n${N}_tailpg$T.kernels.json"))
nc = [k for k in b if "nccl" in k["name"].lower()]
ad = [k for k in b if "multi_tensor_apply" in k["name"] and k["t_us"] > nc[-3]["t_us"]]
print("last 3 nccl:", [(round(k["t_us"] / 1e3, 3), k["dur_us"]) for k in nc[-3:]], "first adam kernel at", round(ad[0]["t_us"] / 1e3, 3), "last kernel ends", round(max(k["t_us"] + k["dur_us"] for k in b) / 1e3, 3))
PY
done
timeout 600 python -m pytest tests/test_parallel_gpu.py tests/test_train_gpu.py -m gpu -q -s > gpurun_out/tests_parallel.log 2>&1; echo "tests rc=$?"; grep -E "passed|failed|skipped|Error|assert" gpurun_out/tests_parallel.log | tail -n 6
