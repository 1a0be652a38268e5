#!/bin/bash
# round 2, call 8: defaults = merged launches + skip convs on a third stream + 2 batches in flight; per-launch table
O=gpurun_out/r2h
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-train --no-eager --no-cpu-baseline --dump-launches $O/launches.txt > $O/bench_default.json 2> $O/bench_default.err
tail -n 4 $O/pytest_gpu.log | cut -c1-200
python - <<PY
import json
d=json.load(open("$O/bench_default.json"))
print({k:round(d[k],3) for k in ("value","ms_per_step")}, "e2e", round(d["e2e"]["value"]), "single", d["single_stream"] and round(d["single_stream"]["ms_per_step"],3), {k:round(d["roofline"][k],4) for k in ("frac","kernel_s_per_step")}, d["gpu_launches_per_step"])
PY
cat $O/launches.txt
