#!/usr/bin/env bash
# N=2 at HEAD: bench line + the reference arm launched the way the driver launches it
set -u
mkdir -p gpurun_out
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench n1 rc=$?"
run 29701 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench n2 rc=$?"
run 29702 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > gpurun_out/bench_ref_n2.json 2> gpurun_out/bench_ref_n2.err; echo "reference arm n2 rc=$?"; wc -l gpurun_out/bench_ref_n2.json; tail -c 300 gpurun_out/bench_ref_n2.json
python - <<PY
import json
for n in (1, 2):
    d = json.loads(open(f"gpurun_out/bench_n{n}.json").read().strip().splitlines()[-1])
    print(n, "value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1), "per-gpu", round(d["value"] / n, 1))
PY
