#!/bin/bash
# round 2, call 53: what do the shift loads (broadcast LDS.128) cost in the conv epilogue?  experiment build without them
O=gpurun_out/r2as
mkdir -p $O
cp dynmm_b200/libdynmm_b200.so /tmp/new.so
for which in base noshift; do
  if [ $which = base ]; then cp /tmp/new.so dynmm_b200/libdynmm_b200.so; else cp tools/bin/libdynmm_noshift.so dynmm_b200/libdynmm_b200.so; fi
  for prec in bf16 f32x3; do
    timeout 600 python bench.py --precision $prec --no-modality --no-cpu-baseline --no-train --no-eager --steps 200 --warmup 10 --dump-launches $O/launches_${which}_${prec}.txt > $O/b_${which}_${prec}.json 2> $O/b_${which}_${prec}.err
    python - <<PY
import json
try:
    d=json.load(open("$O/b_${which}_${prec}.json"))
    print("$which $prec", round(d["value"]), round(d["ms_per_step"],4), "single", d["single_stream"] and round(d["single_stream"]["ms_per_step"],4), "frac", round(d["roofline"]["frac"],4), "kernel_s", d["roofline"]["kernel_s_per_step"])
except Exception as e:
    print("ERR $which $prec", e); print(open("$O/b_${which}_${prec}.err").read()[-1500:])
PY
  done
done
cp /tmp/new.so dynmm_b200/libdynmm_b200.so
