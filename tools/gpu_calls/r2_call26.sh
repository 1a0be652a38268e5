#!/bin/bash
# round 2, call 26: cheap knob sweeps on the f32x3 headline (in-flight depth, TMA L2 promotion)
O=gpurun_out/r2z
mkdir -p $O
B="python bench.py --steps 20 --warmup 5 --no-train --no-eager --no-cpu-baseline --no-modality"
for f in 1 3 4; do timeout 300 $B --in-flight $f > $O/bench_if$f.json 2> $O/bench_if$f.err; done
DYNMM_TMA_L2=256 timeout 300 $B > $O/bench_l2_256.json 2> $O/bench_l2_256.err
DYNMM_TMA_L2=64 timeout 300 $B > $O/bench_l2_64.json 2> $O/bench_l2_64.err
DYNMM_PDL=0 timeout 300 $B > $O/bench_nopdl.json 2> $O/bench_nopdl.err
python - <<PY
import json,glob
for f in sorted(glob.glob("$O/bench_*.json")):
    try:
        d=json.load(open(f))
        print(f.split("/")[-1], d["dtype"], round(d["value"]), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), "| bf16", round(d["bf16"]["value"]), round(d["bf16"]["ms_per_step"],3))
    except Exception as e:
        print(f,"ERR",e)
PY
