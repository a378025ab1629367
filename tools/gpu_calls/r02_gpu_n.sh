#!/usr/bin/env bash
# N ranks on one box: the 2-rank NCCL test, then bench at N=1 and N=$1 on the same box
set -u
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -n 8
timeout 600 python -m pytest tests/test_parallel_gpu.py -m gpu -q -s > gpurun_out/tests_parallel.log 2>&1; echo "tests_parallel rc=$?"; grep -E "passed|failed|skipped|Error|assert" gpurun_out/tests_parallel.log | tail -n 8
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench n1 rc=$?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench n$N rc=$?"; tail -n 5 gpurun_out/bench_n$N.err
python - <<PY
import json
for n in (1, $N):
    try:
        d = json.loads(open(f"gpurun_out/bench_n{n}.json").read().strip().splitlines()[-1])
        print(n, "value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1), "per-gpu", round(d["value"] / n, 1))
    except Exception as e:
        print(n, "no line", e)
PY
