#!/bin/bash
# round 2, call 61: channels_last stem in the bf16 training graph -- A/B of the kernel breakdown, backward tests, train leg
O=gpurun_out/r2az
mkdir -p $O
DYNMM_TRAIN_STEM=nchw timeout 300 python tools/train_profile.py > $O/train_profile_nchw.txt 2>&1
timeout 300 python tools/train_profile.py > $O/train_profile_nhwc.txt 2>&1
head -2 $O/train_profile_nchw.txt | cut -c1-150; head -2 $O/train_profile_nhwc.txt | cut -c1-150
timeout 900 python -m pytest tests -m gpu -q -x -k "backward or train or loss or dp" > $O/pytest_train.log 2>&1; grep -E "passed|failed|FAILED|Error" $O/pytest_train.log | tail -3 | cut -c1-250
for v in nchw nhwc; do
  if [ $v = nchw ]; then export DYNMM_TRAIN_STEM=nchw; else unset DYNMM_TRAIN_STEM; fi
  timeout 600 python bench.py --no-modality --no-cpu-baseline --no-eager --steps 20 --warmup 5 --min-seconds 0 > $O/b_$v.json 2> $O/b_$v.err
  python - <<PY
import json
try:
    d=json.load(open("$O/b_$v.json")); print("$v train", d["train"]["weak_32_per_gpu"]["ms_step"], d["train"]["weak_32_per_gpu"]["loss"])
except Exception as e:
    print("ERR $v", e); print(open("$O/b_$v.err").read()[-1500:])
PY
done
