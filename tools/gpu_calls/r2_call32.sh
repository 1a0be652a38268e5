#!/bin/bash
# round 2, call 32: all four upsampling modes on the engine, 37-class fallback; whole suite
O=gpurun_out/r2af
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-train --no-eager --no-cpu-baseline --no-modality > $O/bench_quick.json 2> $O/bench_quick.err
tail -n 25 $O/pytest_gpu.log | cut -c1-220
python - <<PY
import json
try:
    d=json.load(open("$O/bench_quick.json"))
    print(d["dtype"],{k:round(d[k],3) for k in ("value","ms_per_step")}, "e2e", round(d["e2e"]["value"]), "| bf16", round(d["bf16"]["value"]), round(d["bf16"]["ms_per_step"],3))
except Exception as e:
    print("ERR",e); print(open("$O/bench_quick.err").read()[-1500:])
PY
