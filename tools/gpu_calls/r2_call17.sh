#!/bin/bash
# round 2, call 17: fp32-grade engine mode (f32x3) first run; whole GPU suite; bench in both precisions
O=gpurun_out/r2q
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_f32x3.py -m gpu -q -s > $O/pytest_f32x3.log 2>&1; echo "pytest exit $?" >> $O/pytest_f32x3.log
timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_gpu_f32x3.py > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-train --no-eager --no-cpu-baseline > $O/bench_bf16.json 2> $O/bench_bf16.err
timeout 400 python bench.py --steps 10 --warmup 3 --no-train --no-eager --no-cpu-baseline --precision f32x3 --dump-launches $O/launches_f32x3.txt > $O/bench_f32x3.json 2> $O/bench_f32x3.err
DYNMM_CONV_SMALL=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-train --no-eager --no-cpu-baseline > $O/bench_small.json 2> $O/bench_small.err
tail -n 30 $O/pytest_f32x3.log | cut -c1-250
tail -n 8 $O/pytest_gpu.log | cut -c1-250
python - <<PY
import json
for n in ("bf16","f32x3","small"):
    try:
        d=json.load(open("$O/bench_%s.json"%n))
        print(n,{k:round(d[k],3) for k in ("value","ms_per_step")}, "e2e", round(d["e2e"]["value"]), "single", d["single_stream"] and round(d["single_stream"]["ms_per_step"],3), {k:round(d["roofline"][k],4) for k in ("frac","kernel_s_per_step")}, d["gpu_launches_per_step"])
    except Exception as e:
        print(n,"ERR",e); print(open("$O/bench_%s.err"%n).read()[-1500:])
PY
