#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parallel_gpu.py -m gpu -q -s > gpurun_out/tests_parallel.log 2>&1; echo "tests_parallel rc=$?"; grep -E "passed|failed|\[2-rank" gpurun_out/tests_parallel.log | tail -n 3 | cut -c1-250
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29741 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench n2 rc=$?"
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench n1 rc=$?"
python - <<PY
import json
for n in (1, 2):
    d = json.loads(open(f"gpurun_out/bench_n{n}.json").read().strip().splitlines()[-1])
    print(n, "value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1), "per-gpu", round(d["value"] / n, 1))
PY
grep -c "align2\|align1" gpurun_out/bench_n2.err
