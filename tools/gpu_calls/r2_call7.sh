#!/bin/bash
# round 2, call 7: batches in flight (streams x captured-graph instances) x merged launches
O=gpurun_out/r2g
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_fusion.py tests/test_gpu_kernels.py -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.log
for cfg in "1 0" "2 0" "2 1" "3 0" "3 1" "4 1"; do
  set -- $cfg
  DYNMM_MERGE=$2 timeout 300 python bench.py --steps 20 --warmup 5 --in-flight $1 --no-train --no-eager --no-cpu-baseline > $O/bench_if$1_m$2.json 2> $O/bench_if$1_m$2.err
done
DYNMM_MERGE=1 DYNMM_CONV_DUAL=0 timeout 300 python bench.py --steps 20 --warmup 5 --in-flight 2 --no-train --no-eager --no-cpu-baseline > $O/bench_if2_m1_nodual.json 2> $O/bench_if2_m1_nodual.err
tail -n 4 $O/pytest_gpu.log | cut -c1-200
for f in $O/bench_if*.json; do python - <<PY
import json
try:
    d=json.load(open("$f"))
    print("$f".split("/")[-1], {k:round(d[k],3) for k in ("value","ms_per_step")}, "e2e", round(d["e2e"]["value"]), "single", d["single_stream"] and round(d["single_stream"]["ms_per_step"],3), {k:round(d["roofline"][k],4) for k in ("frac","kernel_s_per_step")})
except Exception as e:
    print("$f", "no json", e)
PY
done
tail -c 1000 $O/bench_if2_m0.err
