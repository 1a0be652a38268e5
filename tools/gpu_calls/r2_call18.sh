#!/bin/bash
# round 2, call 18: f32x3 with TMA-store epilogue; whole suite; the default bench line (f32x3 headline + bf16 block)
O=gpurun_out/r2r
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_f32x3.py -m gpu -q -s > $O/pytest_f32x3.log 2>&1; echo "pytest exit $?" >> $O/pytest_f32x3.log
timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_gpu_f32x3.py > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1
( time timeout 900 python bench.py --dump-launches $O/launches_f32x3.txt > $O/bench_default.json 2> $O/bench_default.err ) 2> $O/bench_time.txt
tail -n 6 $O/pytest_f32x3.log | cut -c1-250
tail -n 4 $O/pytest_gpu.log | cut -c1-250
tail -n 2 $O/smoke.log; cat $O/bench_time.txt | tail -4
python - <<PY
import json
try:
    d=json.load(open("$O/bench_default.json"))
    print(d["dtype"],{k:round(d[k],3) for k in ("value","ms_per_step")}, "e2e", round(d["e2e"]["value"]), "single", d["single_stream"] and round(d["single_stream"]["ms_per_step"],3), {k:round(d["roofline"][k],4) for k in ("frac","kernel_s_per_step")}, d["gpu_launches_per_step"])
    o=d["bf16"]; print("bf16", round(o["value"]), round(o["ms_per_step"],3), "e2e", round(o["e2e"]["value"]), "frac", round(o["roofline"]["frac"],4))
    print("eager", {k:(round(v,2) if isinstance(v,float) else v) for k,v in d["gpu_eager_baseline"].items() if "over" in k})
    print("train", d["train"]["weak_32_per_gpu"]["ms_step"], "cpu", d["cpu_baseline"])
except Exception as e:
    print("ERR",e); print(open("$O/bench_default.err").read()[-2500:])
PY
