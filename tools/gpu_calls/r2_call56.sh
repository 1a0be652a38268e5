#!/bin/bash
# round 2, call 56: ncu evidence on the final tree (both precisions + bench launch list), three batches in flight
O=gpurun_out/r2av
mkdir -p $O
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active
PRECISION=f32x3 timeout 400 ncu --profile-from-start off --clock-control none --metrics $M --csv --log-file $O/r2_f32x3_step_metrics.csv python tools/one_step.py > $O/one_step_f32x3.log 2>&1
PRECISION=bf16 timeout 400 ncu --profile-from-start off --clock-control none --metrics $M --csv --log-file $O/r2_bf16_step_metrics.csv python tools/one_step.py > $O/one_step_bf16.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r2_bench_launches.csv python bench.py --steps 2 --warmup 3 --min-seconds 0 --no-train --no-eager --no-cpu-baseline --no-modality > $O/bench_under_ncu.log 2>&1
tail -1 $O/one_step_f32x3.log; tail -1 $O/one_step_bf16.log
for fl in 2 3; do
for prec in f32x3 bf16; do
  timeout 600 python bench.py --precision $prec --in-flight $fl --no-modality --no-cpu-baseline --no-train --no-eager --steps 200 --warmup 10 > $O/b_${prec}_$fl.json 2> $O/b_${prec}_$fl.err
  python - <<PY
import json
try:
    d=json.load(open("$O/b_${prec}_$fl.json"))
    print("$prec in-flight $fl", round(d["value"]), round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"]))
except Exception as e:
    print("ERR $prec", e); print(open("$O/b_${prec}_$fl.err").read()[-1500:])
PY
done
done
