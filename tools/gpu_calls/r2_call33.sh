#!/bin/bash
# round 2, call 33: chain kernel triggers its dependent launch at the end -- chain tests + bench (bf16 headline run too)
O=gpurun_out/r2ag
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_chain.py tests/test_gpu_fusion.py -m gpu -q > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-train --no-eager --no-cpu-baseline --no-modality --precision bf16 > $O/bench_bf16.json 2> $O/bench_bf16.err
timeout 600 python bench.py --steps 20 --warmup 5 --no-train --no-eager --no-cpu-baseline --no-modality --precision bf16 > $O/bench_bf16_b.json 2> $O/bench_bf16_b.err
tail -n 3 $O/pytest.log | cut -c1-200
python - <<PY
import json
for n in ("bf16","bf16_b"):
    try:
        d=json.load(open("$O/bench_%s.json"%n))
        print(d["dtype"],{k:round(d[k],3) for k in ("value","ms_per_step")}, "e2e", round(d["e2e"]["value"]), "single", round(d["single_stream"]["ms_per_step"],3), "| f32x3", round(d["f32x3"]["value"]), round(d["f32x3"]["ms_per_step"],3))
    except Exception as e:
        print("ERR",e); print(open("$O/bench_%s.err"%n).read()[-1500:])
PY
