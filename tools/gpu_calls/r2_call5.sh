#!/bin/bash
# round 2, call 5: merged RGB+depth launches (dynmm_conv_igemm_fwd2)
O=gpurun_out/r2e
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.log
DYNMM_MERGE=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-train --no-eager --no-cpu-baseline > $O/bench_merge.json 2> $O/bench_merge.err
DYNMM_MERGE=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-train --no-eager --no-cpu-baseline > $O/bench_nomerge.json 2> $O/bench_nomerge.err
DYNMM_MERGE=1 DYNMM_CONV_DUAL=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-train --no-eager --no-cpu-baseline > $O/bench_merge_nodual.json 2> $O/bench_merge_nodual.err
DYNMM_MERGE=1 DYNMM_CONV_DUAL=2 timeout 300 python bench.py --steps 20 --warmup 5 --no-train --no-eager --no-cpu-baseline > $O/bench_merge_dual2.json 2> $O/bench_merge_dual2.err
tail -n 12 $O/pytest_gpu.log | cut -c1-300
for f in merge nomerge merge_nodual merge_dual2; do echo $f; python - <<PY
import json
try:
    d=json.load(open("$O/bench_$f.json"))
    print({k:d[k] for k in ("value","ms_per_step","gpu_launches_per_step")}, {k:d["roofline"][k] for k in ("frac","kernel_s_per_step","launches_per_step")})
except Exception as e:
    print("no json", e)
PY
done
tail -c 1500 $O/bench_merge.err
