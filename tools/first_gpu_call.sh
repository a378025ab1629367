#!/usr/bin/env bash
# One gpurun call that validates everything written after round 1's GPU budget ran out, in the order that matters:
#   /usr/local/graft/bin/gpurun --timeout 2400 -- bash tools/first_gpu_call.sh
# Every step writes its own log under gpurun_out/; a failing step does not stop the later ones.
set -u
mkdir -p gpurun_out
run() { local name=$1; shift; echo "== $name"; timeout 1200 "$@" > "gpurun_out/$name.log" 2>&1; echo "   rc=$? ($(tail -n 1 "gpurun_out/$name.log" | cut -c1-150))"; }

python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
# 1. the product path: full GPU suite (the model-family tests in test_variants_gpu.py are new), smoke, default bench
run tests_gpu            python -m pytest tests -m gpu -q -x --deselect tests/test_variants_gpu.py
run tests_variants       python -m pytest tests/test_variants_gpu.py -m gpu -q
run smoke                python -c "import __graft_entry__ as g; g.smoke()"
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
# 2. the opt-in kernels: a staged probe that logs before every launch (a hang still tells where), then the full parity
#    file, then timings alone and inside the step
PVQA_ATTN_BWD_LEAN=1 run probe python tools/quick_v2_probe.py
run timing python tools/quick_v2_timing.py
PVQA_TEST_ATTN_V2=1 run tests_attn_v2 python -m pytest tests/test_attn_v2_gpu.py -m gpu -q
run kbench_attn_v1       python tools/kbench.py attn
PVQA_ATTN_FWD_V2=1 run kbench_attn_v2 python tools/kbench.py attn
PVQA_ATTN_FWD_V3=1 run kbench_attn_v3 python tools/kbench.py attn
PVQA_ATTN_FWD_V3=1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_attn_v3.json 2> gpurun_out/bench_attn_v3.err
PVQA_ATTN_FWD_V2=1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_attn_v2.json 2> gpurun_out/bench_attn_v2.err
# 2b. the lean backward variant (SCP code compiled out of non-SaL launches): same tests in a process that opts in
PVQA_ATTN_BWD_LEAN=1 run tests_bwd_lean python -m pytest tests/test_attn_gpu.py tests/test_model_gpu.py -m gpu -q
PVQA_ATTN_BWD_LEAN=1 run kbench_attn_lean python tools/kbench.py attn
PVQA_ATTN_BWD_LEAN=1 PVQA_ATTN_FWD_V2=1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v2_lean.json 2> gpurun_out/bench_v2_lean.err
PVQA_ATTN_BWD_LEAN=1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_lean.json 2> gpurun_out/bench_lean.err
# 3. sibling workloads and the eager-PyTorch comparison (not yet measured)
python bench.py --workload phonoprestu --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_prestu224.json 2> gpurun_out/bench_prestu224.err
python bench.py --workload phonoprestu --image 384 --batch 32 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_prestu384.json 2> gpurun_out/bench_prestu384.err
python bench.py --workload phonosal --batch 32 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_sal.json 2> gpurun_out/bench_sal.err
python bench.py --impl eager-gpu --batch 64 --steps 5 --warmup 2 > gpurun_out/bench_eager_gpu.json 2> gpurun_out/bench_eager_gpu.err

PVQA_ATTN_FWD_V2=1 run attn_trace_v2 bash -c "python tools/attn_trace.py build && python tools/attn_trace.py 64"
ls -la gpurun_out | tail -n 24
PVQA_ATTN_FWD_V3=1 run attn_trace_v3 python tools/attn_trace.py 64
# compact summary (also kept as gpurun_out/summary.txt): last line of every log, headline numbers of every bench line
python - <<'PY' | tee gpurun_out/summary.txt
import glob, json, os
for f in sorted(glob.glob("gpurun_out/*.log")):
    lines = [l.rstrip() for l in open(f, errors="ignore") if l.strip()]
    print(f"{os.path.basename(f):28s} {lines[-1][:150] if lines else '(empty)'}")
for f in sorted(glob.glob("gpurun_out/bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        extra = {k: round(v["samples_per_s"], 1) for k, v in d.get("variants", {}).items()}
        print(f"{os.path.basename(f):28s} {d.get('metric', '?')}: value {d.get('value', 0):.1f} "
              f"ms/step {d.get('ms_per_step', 0) or 0:.2f} e2e {(d.get('e2e') or {}).get('value', 0):.1f} {extra or ''}")
    except Exception as e:
        err = f.replace(".json", ".err")
        tail = open(err, errors="ignore").read().strip().splitlines()[-1][:150] if os.path.exists(err) else ""
        print(f"{os.path.basename(f):28s} no JSON line ({type(e).__name__}); stderr: {tail}")
for f in ("gpurun_out/quick_v2_probe.log", "gpurun_out/quick_v2_timing.log"):
    if os.path.exists(f):
        print("----", f)
        print(open(f).read())
PY
