#!/bin/bash
# round 2, 8-GPU call at the end of the round: the driver's scaling invocation of bench.py on the final tree
O=gpurun_out/r2bd
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
( time timeout 400 $TR --master-port 29533 bench.py --gpus 8 > $O/bench_n8.json 2> $O/bench_n8.err ) 2> $O/bench_n8_time.txt
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
