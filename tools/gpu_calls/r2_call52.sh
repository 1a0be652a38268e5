#!/bin/bash
# round 2, call 52: stem with parameter BN + PDL after the pre-pass: suite, stem timing, quick bench lines
O=gpurun_out/r2at
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.log
grep -E "passed|failed|FAILED|Error" $O/pytest_gpu.log | tail -4 | cut -c1-250
timeout 300 python tools/stem_bench.py 2>&1 | tail -8 | tee $O/stem_bench.txt
for prec in bf16 f32x3; do
  timeout 600 python bench.py --precision $prec --no-modality --no-cpu-baseline --no-train --no-eager --steps 200 --warmup 10 > $O/b_${prec}.json 2> $O/b_${prec}.err
  python - <<PY
import json
try:
    d=json.load(open("$O/b_${prec}.json"))
    print("$prec", round(d["value"]), round(d["ms_per_step"],4), "single", d["single_stream"] and round(d["single_stream"]["ms_per_step"],4), "frac", round(d["roofline"]["frac"],4))
except Exception as e:
    print("ERR $prec", e); print(open("$O/b_${prec}.err").read()[-1500:])
PY
done
