"""profiles/traffic.json from an `ncu --set full` capture of tools/ncu_one_layer.py (one-layer PhonemeLaTr at the bench
shape: every kernel family of libpvqa_sm100.so at exactly the shapes of the 12-layer step).
    ncu -i rep.ncu-rep --page raw --csv > raw.csv ;  python tools/ncu_traffic.py raw.csv profiles/traffic.json
traffic = dram__bytes_read.sum + dram__bytes_write.sum per launch, keyed by the kernel names bench.py reports.  Kernels
bench.py times as one entry but that launch several kernels (attention backward = prep + main) are summed; kernels used
at two shapes inside the step (encoder 20928 x 3072 and decoder 8128 x 2048 feed-forward) are averaged with the step's
launch mix (12 encoder : 4 decoder layers)."""
import csv
import json
import re
import sys

src, dst = sys.argv[1], sys.argv[2]
rows = list(csv.reader(open(src)))
hdr = rows[0]
idx = {h: i for i, h in enumerate(hdr)}
unit_scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
ur, uw = rows[1][idx["dram__bytes_read.sum"]], rows[1][idx["dram__bytes_write.sum"]]
launches = []
for r in rows[2:]:
    if len(r) < len(hdr) or not r[0].isdigit():
        continue
    name = r[idx["Kernel Name"]]
    by = float(r[idx["dram__bytes_read.sum"]].replace(",", "")) * unit_scale[ur] + \
        float(r[idx["dram__bytes_write.sum"]].replace(",", "")) * unit_scale[uw]
    launches.append((name, r[idx["Grid Size"]], float(r[idx["gpu__time_duration.sum"]].replace(",", "")), by))


def pick(pattern, grid=None):
    return [(t, b) for n, g, t, b in launches if re.search(pattern, n) and (grid is None or g.startswith(f"({grid},"))]


def mean(xs):
    xs = list(xs)
    return sum(xs) / len(xs) if xs else None


def mix(enc, dec, n_enc=12, n_dec=4):
    return (n_enc * mean(b for _, b in enc) + n_dec * mean(b for _, b in dec)) / (n_enc + n_dec)


out = {"_comment": "ncu dram__bytes_read.sum + dram__bytes_write.sum per launch (bytes), one-layer PhonemeLaTr at the bench "
                   "shape B=64 d=768 H=12 S=327 T=127 p=0.1 (tools/ncu_one_layer.py, profiles/r02_ncu_kernels.txt); cold L2, "
                   "launches serialised by the profiler"}
prep = pick(r"attn_bwd_prep_kernel")
# launch order of the backward: cross (Sq=127,Sk=327), decoder self (127), encoder self (327); prep kernels in the same order
attn = {"attn_fwd[Sq=197,Sk=197]": pick(r"attn_fwd_kernel<0, 0, 0, 0>"), "attn_fwd[Sq=327,Sk=327]": pick(r"attn_fwd_kernel<1, 1, 0, 0>"),
        "attn_fwd[Sq=127,Sk=127]": pick(r"attn_fwd_kernel<0, 1, 1, 0>"), "attn_fwd[Sq=127,Sk=327]": pick(r"attn_fwd_kernel<0, 1, 0, 0>")}
for k, v in attn.items():
    out[k] = mean(b for _, b in v)
for key, pat, pi in (("attn_bwd[Sq=127,Sk=327]", r"attn_bwd_kernel<0, 1, 0, 0>", 0), ("attn_bwd[Sq=127,Sk=127]", r"attn_bwd_kernel<0, 1, 1, 0>", 1),
                     ("attn_bwd[Sq=327,Sk=327]", r"attn_bwd_kernel<1, 1, 0, 0>", 2)):
    out[key] = mean(b for _, b in pick(pat)) + prep[pi][1]
out["add_ln_lp"] = mean(b for _, b in pick(r"add_dropout_ln_fwd_kernel<__nv_bfloat16, __nv_bfloat16, __nv_bfloat16"))
out["add_dropout_rms_fwd"] = mean(b for _, b in pick(r"add_dropout_ln_fwd_kernel<float, __nv_bfloat16, __nv_bfloat16", 1184))
out["add_dropout_ln_fwd"] = mean(b for _, b in pick(r"add_dropout_ln_fwd_kernel<float, __nv_bfloat16, __nv_bfloat16", 1016))
out["add_dropout_rms_bwd"] = mean(b for _, b in pick(r"add_dropout_ln_bwd_kernel<__nv_bfloat16, __nv_bfloat16, 3, 0>"))
out["add_dropout_ln_bwd"] = mean(b for _, b in pick(r"add_dropout_ln_bwd_kernel<__nv_bfloat16, __nv_bfloat16, 3, 1>"))
rf, rb, cr = pick(r"relu_dropout_fwd_kernel"), pick(r"relu_dropout_bwd_kernel"), pick(r"cast_rows_kernel")
big = lambda v: [x for x in v if x[1] == max(b for _, b in v)]       # noqa: E731  encoder-shape launch
small = lambda v: [x for x in v if x[1] == min(b for _, b in v)]     # noqa: E731  decoder-shape launch
out["relu_dropout_fwd"] = mix(big(rf), small(rf))
out["relu_dropout_bwd"] = mix(big(rb), small(rb))
out["cast_rows"] = mix(big(cr), small(cr))
out["col_sum"] = mean(b for _, b in pick(r"col_sum_kernel"))
for key, pat in (("rms_norm_fwd", r"rms_norm_fwd_kernel"), ("rms_norm_bwd", r"rms_norm_bwd_kernel"), ("embed_mm_fwd", r"embed_mm_fwd_kernel"),
                 ("embed_mm_bwd", r"embed_mm_bwd_kernel"), ("embed_tgt_fwd", r"embed_tgt_fwd"), ("embed_tgt_bwd", r"embed_tgt_bwd_kernel"),
                 ("phoneme_head_fused_fwd", r"phoneme_head_tc_kernel"), ("phoneme_head_ce_bwd", r"phoneme_head_mma_kernel<1>")):
    v = pick(pat)
    if v:
        out[key] = mean(b for _, b in v)
out = {k: (v if isinstance(v, str) else int(round(v))) for k, v in out.items() if v is not None}
json.dump(out, open(dst, "w"), indent=1)
print(json.dumps(out, indent=1))
