#!/bin/bash
# round 2, call 19: split as a compile-time kernel mode -- regression run + both bench precisions
O=gpurun_out/r2s
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-train --no-eager --no-cpu-baseline --dump-launches $O/launches_f32x3.txt > $O/bench_quick.json 2> $O/bench_quick.err
timeout 600 python bench.py --steps 20 --warmup 5 --no-train --no-eager --no-cpu-baseline --precision bf16 --dump-launches $O/launches_bf16.txt > $O/bench_quick_bf16.json 2> $O/bench_quick_bf16.err
tail -n 4 $O/pytest_gpu.log | cut -c1-250
python - <<PY
import json
for n in ("quick","quick_bf16"):
    try:
        d=json.load(open("$O/bench_%s.json"%n))
        o=[k for k in ("bf16","f32x3") if k in d and d[k]][0]
        print(d["dtype"],{k:round(d[k],3) for k in ("value","ms_per_step")}, "e2e", round(d["e2e"]["value"]), "single", d["single_stream"] and round(d["single_stream"]["ms_per_step"],3), {k:round(d["roofline"][k],4) for k in ("frac","kernel_s_per_step")}, d["gpu_launches_per_step"], "| other", o, round(d[o]["value"]), round(d[o]["ms_per_step"],3), round(d[o]["roofline"]["frac"],4))
    except Exception as e:
        print(n,"ERR",e); print(open("$O/bench_%s.err"%n).read()[-1500:])
PY
