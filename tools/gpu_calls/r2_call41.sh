#!/bin/bash
# round 2, call 41: swish / h-swish models on the engine; whole suite + bench regression check
O=gpurun_out/r2an
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-train --no-eager --no-cpu-baseline --no-modality > $O/bench_quick.json 2> $O/bench_quick.err
grep -E "passed|failed|^E  |FAILED" $O/pytest_gpu.log | tail -10 | cut -c1-250
python - <<PY
import json
try:
    d=json.load(open("$O/bench_quick.json"))
    print(d["dtype"],{k:round(d[k],3) for k in ("value","ms_per_step")}, "e2e", round(d["e2e"]["value"]), "single", round(d["single_stream"]["ms_per_step"],3), "| bf16", round(d["bf16"]["value"]), round(d["bf16"]["ms_per_step"],3))
except Exception as e:
    print("ERR",e); print(open("$O/bench_quick.err").read()[-1500:])
PY
