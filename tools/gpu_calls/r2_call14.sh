#!/bin/bash
# round 2, call 14: chain kernel in the engine (stage 3), full GPU suite subset + bench
O=gpurun_out/r2n
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_chain.py tests/test_gpu_kernels.py tests/test_gpu_fusion.py tests/test_gpu_pair.py tests/test_gpu_program.py -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.log
timeout 300 python tools/chain_bench.py > $O/chain_bench.txt 2>&1
timeout 300 python bench.py --steps 20 --warmup 5 --no-train --no-eager --no-cpu-baseline --dump-launches $O/launches.txt > $O/bench_default.json 2> $O/bench_default.err
DYNMM_CHAIN_STAGES=1,2 timeout 300 python bench.py --steps 20 --warmup 5 --no-train --no-eager --no-cpu-baseline --dump-launches $O/launches_s12.txt > $O/bench_s12.json 2> $O/bench_s12.err
DYNMM_CHAIN=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-train --no-eager --no-cpu-baseline > $O/bench_nochain.json 2> $O/bench_nochain.err
tail -n 4 $O/pytest_gpu.log | cut -c1-300
cat $O/chain_bench.txt | tail -6
python - <<PY
import json
for n in ("default","s12","nochain"):
    try:
        d=json.load(open("$O/bench_%s.json"%n))
        print(n,{k:round(d[k],3) for k in ("value","ms_per_step")}, "e2e", round(d["e2e"]["value"]), "single", d["single_stream"] and round(d["single_stream"]["ms_per_step"],3), {k:round(d["roofline"][k],4) for k in ("frac","kernel_s_per_step")}, d["gpu_launches_per_step"])
    except Exception as e:
        print(n,"ERR",e); print(open("$O/bench_%s.err"%n).read()[-1500:])
PY
