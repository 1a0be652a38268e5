#!/bin/bash
# round 2, call 57: wave-aware strip length of the decoder up-sampling kernel, A/B against the previous library
O=gpurun_out/r2aw
mkdir -p $O
cp dynmm_b200/libdynmm_b200.so /tmp/new.so
for which in prev new; do
  if [ $which = prev ]; then cp tools/bin/libdynmm_prev.so dynmm_b200/libdynmm_b200.so; else cp /tmp/new.so dynmm_b200/libdynmm_b200.so; fi
  echo "== $which" | tee -a $O/upsample_bench.txt
  timeout 300 python tools/upsample_bench.py 2>&1 | tail -8 | tee -a $O/upsample_bench.txt
done
cp /tmp/new.so dynmm_b200/libdynmm_b200.so
timeout 900 python -m pytest tests -m gpu -q -x -k "upsample or f32x3 or fusion or golden" > $O/pytest.log 2>&1; grep -E "passed|failed|FAILED|Error" $O/pytest.log | tail -3 | cut -c1-250
for prec in f32x3 bf16; do
  timeout 600 python bench.py --precision $prec --no-modality --no-cpu-baseline --no-train --no-eager --steps 200 --warmup 10 > $O/b_${prec}.json 2> $O/b_${prec}.err
  python - <<PY
import json
try:
    d=json.load(open("$O/b_${prec}.json"))
    print("$prec", round(d["value"]), round(d["ms_per_step"],4), "single", d["single_stream"] and round(d["single_stream"]["ms_per_step"],4))
except Exception as e:
    print("ERR $prec", e); print(open("$O/b_${prec}.err").read()[-1500:])
PY
done
