#!/usr/bin/env bash
# ncu --set full of the four attention-forward launches of one step, final tile rule (one masked remainder tile)
set -u
mkdir -p gpurun_out
timeout 300 python tools/ncu_one_layer.py > gpurun_out/ncu_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:^attn_fwd -c 4 -f -o /tmp/r02_attn_fwd python tools/ncu_one_layer.py > gpurun_out/ncu_run.log 2>&1; echo "ncu rc=$?"; tail -n 2 gpurun_out/ncu_run.log
ncu -i /tmp/r02_attn_fwd.ncu-rep --page raw --csv > gpurun_out/r02_attn_fwd_raw.csv 2> gpurun_out/ncu_export.err; echo "export rc=$?"
python tools/ncu_summary.py < gpurun_out/r02_attn_fwd_raw.csv > gpurun_out/r02_ncu_attn_fwd_final.txt; grep -E "^==|gpu__time_duration|dram__bytes|tensor_cycles|issue_active" gpurun_out/r02_ncu_attn_fwd_final.txt
cp /tmp/r02_attn_fwd.ncu-rep gpurun_out/ 2>/dev/null; ls -la gpurun_out/*.ncu-rep
