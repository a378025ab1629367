#!/usr/bin/env bash
# new training tests, every workload with the current kernels, the PyTorch-eager incumbent of each, ncu evidence
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_train_gpu.py tests/test_parallel_gpu.py -m gpu -q -x -s > gpurun_out/tests_train.log 2>&1; echo "tests_train rc=$?"; grep -E "passed|failed|gradient quality|Error|assert" gpurun_out/tests_train.log | tail -n 12
for w in phonolatr latr phonosal; do
  timeout 600 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline $( [ $w = phonosal ] && echo "--batch 32" ) > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; echo "bench $w rc=$?"
done
timeout 600 python bench.py --workload phonoprestu --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_prestu224.json 2> gpurun_out/bench_prestu224.err; echo "bench prestu224 rc=$?"
timeout 600 python bench.py --workload phonoprestu --image 384 --batch 32 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_prestu384.json 2> gpurun_out/bench_prestu384.err; echo "bench prestu384 rc=$?"
for w in phonolatr latr phonoprestu phonosal; do
  timeout 600 python bench.py --impl eager-gpu --workload $w --steps 4 --warmup 2 $( [ $w = phonosal ] && echo "--batch 32" ) > gpurun_out/eager_$w.json 2> gpurun_out/eager_$w.err; echo "eager $w rc=$?"
done
python - <<'PY'
import glob, json, os
for f in sorted(glob.glob("gpurun_out/bench_*.json") + glob.glob("gpurun_out/eager_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f"{os.path.basename(f):26s} value {d.get('value', 0):9.1f} ms/step {d.get('ms_per_step') or 0:7.2f} e2e {(d.get('e2e') or {}).get('value', 0):9.1f} {d.get('variants') or ''}")
    except Exception as e:
        err = f.replace(".json", ".err")
        print(os.path.basename(f), "no JSON line;", open(err, errors="ignore").read().strip().splitlines()[-1][:200] if os.path.exists(err) else e)
PY
# ncu: plain run of the same command first, then the capture (one-layer model: every kernel family at bench shape)
python tools/ncu_one_layer.py > gpurun_out/ncu_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:pvqa -c 120 -o gpurun_out/r02_kernels python tools/ncu_one_layer.py > gpurun_out/ncu_run.log 2>&1; echo "ncu rc=$?"; tail -n 3 gpurun_out/ncu_run.log
ls -la gpurun_out/*.ncu-rep 2>/dev/null
