#!/bin/bash
# round 2, call 15: chain kernel -- thread-loaded input, 8 vs 16 epilogue warps
O=gpurun_out/r2o
mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_chain.py -m gpu -q > $O/pytest_chain16.log 2>&1; echo "pytest exit $?" >> $O/pytest_chain16.log
DYNMM_CHAIN_EPI=8 timeout 300 python -m pytest tests/test_gpu_chain.py -m gpu -q > $O/pytest_chain8.log 2>&1; echo "pytest exit $?" >> $O/pytest_chain8.log
timeout 300 python tools/chain_bench.py > $O/chain_bench16.txt 2>&1
DYNMM_CHAIN_EPI=8 timeout 300 python tools/chain_bench.py > $O/chain_bench8.txt 2>&1
timeout 300 python tools/chain_trace.py > $O/chain_trace16.txt 2>&1
DYNMM_CHAIN_TRACE_LAYER=0 timeout 300 python tools/chain_trace.py > $O/chain_trace16_l0.txt 2>&1
timeout 300 python bench.py --steps 20 --warmup 5 --no-train --no-eager --no-cpu-baseline --dump-launches $O/launches.txt > $O/bench_default.json 2> $O/bench_default.err
tail -n 3 $O/pytest_chain16.log | cut -c1-300; tail -n 3 $O/pytest_chain8.log | cut -c1-300
echo "-- 16 warps"; tail -6 $O/chain_bench16.txt; echo "-- 8 warps"; tail -6 $O/chain_bench8.txt
grep -A 13 "stage3 rgb+depth(4)" $O/chain_trace16.txt; grep -A 8 "stage3 rgb+depth(4)" $O/chain_trace16_l0.txt
python - <<PY
import json
for n in ("default",):
    try:
        d=json.load(open("$O/bench_%s.json"%n))
        print(n,{k:round(d[k],3) for k in ("value","ms_per_step")}, "e2e", round(d["e2e"]["value"]), "single", d["single_stream"] and round(d["single_stream"]["ms_per_step"],3), {k:round(d["roofline"][k],4) for k in ("frac","kernel_s_per_step")}, d["gpu_launches_per_step"])
    except Exception as e:
        print(n,"ERR",e); print(open("$O/bench_%s.err"%n).read()[-1500:])
PY
