#!/bin/bash
# round 2, call 45 / 66: A/B of a conv epilogue change (new .so in tree, previous one in tools/bin)
O=gpurun_out/r2bb
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.log
grep -E "passed|failed|FAILED" $O/pytest_gpu.log | tail -4 | cut -c1-250
cp dynmm_b200/libdynmm_b200.so /tmp/new.so
for rep in 1 2; do
for which in new prev; do
  if [ $which = prev ]; then cp tools/bin/libdynmm_prev.so dynmm_b200/libdynmm_b200.so; else cp /tmp/new.so dynmm_b200/libdynmm_b200.so; fi
  for prec in bf16 f32x3; do
    timeout 600 python bench.py --precision $prec --no-modality --no-cpu-baseline --no-train --no-eager --steps 200 --warmup 10 > $O/b_${which}_${prec}_$rep.json 2> $O/b_${which}_${prec}_$rep.err
    python - <<PY
import json
try:
    d=json.load(open("$O/b_${which}_${prec}_$rep.json"))
    print("$which $prec $rep", round(d["value"]), round(d["ms_per_step"],4), "single", d["single_stream"] and round(d["single_stream"]["ms_per_step"],4), "frac", round(d["roofline"]["frac"],4))
except Exception as e:
    print("ERR $which $prec", e); print(open("$O/b_${which}_${prec}_$rep.err").read()[-1500:])
PY
  done
done
done
cp /tmp/new.so dynmm_b200/libdynmm_b200.so
