#!/bin/bash
# round 2, call 23: f32x3 residual prefetch + strip up-sampling; regression
O=gpurun_out/r2w
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_f32x3.py tests/test_gpu_fusion.py tests/test_gpu_kernels.py -m gpu -q -x > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-train --no-eager --no-cpu-baseline --dump-launches $O/launches_f32x3.txt > $O/bench_quick.json 2> $O/bench_quick.err
SPLIT=1 ONLY="s1 1x3 c64 +res" timeout 200 python tools/conv_trace.py > $O/trace_split_s1_res.txt 2>&1
tail -n 3 $O/pytest.log | cut -c1-250
python - <<PY
import json
try:
    d=json.load(open("$O/bench_quick.json"))
    print(d["dtype"],{k:round(d[k],3) for k in ("value","ms_per_step")}, "e2e", round(d["e2e"]["value"]), "single", d["single_stream"] and round(d["single_stream"]["ms_per_step"],3), {k:round(d["roofline"][k],4) for k in ("frac","kernel_s_per_step")}, d["gpu_launches_per_step"], "| bf16", round(d["bf16"]["value"]), round(d["bf16"]["ms_per_step"],3), round(d["bf16"]["roofline"]["frac"],4))
except Exception as e:
    print("ERR",e); print(open("$O/bench_quick.err").read()[-1500:])
PY
head -16 $O/trace_split_s1_res.txt
