#!/bin/bash
# round 2, 8-GPU call: configs[4] robustness sweep on 8 x B200 (both engine precisions), then the driver's scaling invocation
O=gpurun_out/r2x
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29521 tools/noise_sweep.py --batches 24 --precision f32x3 > $O/noise_sweep_8gpu_f32x3.txt 2> $O/noise_f32x3.err
timeout 400 $TR --master-port 29522 tools/noise_sweep.py --batches 24 --precision bf16 > $O/noise_sweep_8gpu_bf16.txt 2> $O/noise_bf16.err
( time timeout 900 $TR --master-port 29523 bench.py --gpus 8 > $O/bench_n8.json 2> $O/bench_n8.err ) 2> $O/bench_n8_time.txt
tail -1 $O/noise_sweep_8gpu_f32x3.txt | cut -c1-600; tail -1 $O/noise_sweep_8gpu_bf16.txt | cut -c1-600
tail -3 $O/bench_n8_time.txt
python - <<PY
import json
try:
    txt=[l for l in open("$O/bench_n8.json") if l.startswith("{")][-1]
    d=json.loads(txt)
    print(d["dtype"], d["n_gpus"], {k:round(d[k],3) for k in ("value","ms_per_step")}, "e2e", round(d["e2e"]["value"]), "ranks", [round(x,3) for x in d["rank_ms_per_step"]], "bf16", round(d["bf16"]["value"]), round(d["bf16"]["ms_per_step"],3))
    print("train", json.dumps(d["train"])[:1200])
except Exception as e:
    print("ERR", e); print(open("$O/bench_n8.err").read()[-3000:])
PY
