#!/bin/bash
# round 2, call 42: A/B of the activation branch in the bf16 epilogue (current library vs the previous commit's)
O=gpurun_out/r2ao
mkdir -p $O
B="python bench.py --steps 20 --warmup 5 --no-train --no-eager --no-cpu-baseline --no-modality --precision bf16"
for i in 1 2; do timeout 300 $B > $O/cur_$i.json 2> $O/cur_$i.err; done
cp dynmm_b200/libdynmm_b200.so /tmp/cur.so; cp dynmm_b200/libdynmm_b200_prev.so.keep dynmm_b200/libdynmm_b200.so
for i in 1 2; do timeout 300 $B > $O/prev_$i.json 2> $O/prev_$i.err; done
cp /tmp/cur.so dynmm_b200/libdynmm_b200.so
for i in 3; do timeout 300 $B > $O/cur_$i.json 2> $O/cur_$i.err; done
python - <<PY
import json,glob
for f in sorted(glob.glob("$O/*.json")):
    try:
        d=json.load(open(f)); print(f.split("/")[-1], d["dtype"], round(d["value"]), round(d["ms_per_step"],3), "single", round(d["single_stream"]["ms_per_step"],3), "kernel_s", round(d["roofline"]["kernel_s_per_step"]*1e3,3), "| f32x3", round(d["f32x3"]["value"]))
    except Exception as e: print(f,"ERR",e)
PY
