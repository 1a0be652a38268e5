#!/bin/bash
# round 2, call 36: final verification -- whole suite, smoke, default bench line, reference arm
O=gpurun_out/r2aj
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1
( time timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err ) 2> $O/bench_time.txt
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $O/bench_reference.json 2> $O/bench_reference.err
grep -E "passed|failed|FAILED" $O/pytest_gpu.log | tail -4 | cut -c1-250; tail -n 1 $O/smoke.log; tail -3 $O/bench_time.txt
python - <<PY
import json
try:
    d=json.load(open("$O/bench_default.json"))
    print(d["dtype"],{k:round(d[k],3) for k in ("value","ms_per_step")}, "e2e", round(d["e2e"]["value"]), "single", d["single_stream"] and round(d["single_stream"]["ms_per_step"],3), {k:round(d["roofline"][k],4) for k in ("frac","kernel_s_per_step")}, d["gpu_launches_per_step"])
    o=d["bf16"]; print("bf16", round(o["value"]), round(o["ms_per_step"],3), "e2e", round(o["e2e"]["value"]), "frac", round(o["roofline"]["frac"],4))
    print("sweep", {k:(round(v["images_per_s"]), round(v["flop_saved_pct"],1)) for k,v in d["branch_sweep"].items() if isinstance(v,dict)})
    print("eager", {k:(round(v,2) if isinstance(v,float) else v) for k,v in d["gpu_eager_baseline"].items() if "over" in k})
    print("train", d["train"]["weak_32_per_gpu"]["ms_step"], "cpu", d["cpu_baseline"]["value"], "ref", json.load(open("$O/bench_reference.json"))["value"])
except Exception as e:
    print("ERR",e); print(open("$O/bench_default.err").read()[-2500:])
PY
