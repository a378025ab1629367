#!/usr/bin/env bash
# N=4: attention-backward granularity 1 vs 4 CTAs per SM under the collectives, then the bench line
set -u
N=4
mkdir -p gpurun_out
P=29660
run() { P=$((P+1)); timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P "$@"; }
for W in 1 4; do
run tools/ddp_timeline.py --waves $W --out gpurun_out/ddp_timeline_n${N}_waves$W > gpurun_out/tl_w$W.log 2>&1; echo "timeline waves $W rc=$?"; head -n 1 gpurun_out/ddp_timeline_n${N}_waves$W.txt | cut -c1-200 || tail -n 5 gpurun_out/tl_w$W.log
grep "attn_bwd_kernel<true" gpurun_out/ddp_timeline_n${N}_waves$W.txt | cut -c1-160
done
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench n1 rc=$?"
run bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench n$N rc=$?"
python - <<PY
import json
for n in (1, $N):
    try:
        d = json.loads(open(f"gpurun_out/bench_n{n}.json").read().strip().splitlines()[-1])
        print(n, "value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1), "per-gpu", round(d["value"] / n, 1), d["config"].get("attn_bwd_waves"))
    except Exception as e:
        print(n, "no line", e)
PY
