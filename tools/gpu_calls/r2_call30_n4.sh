#!/bin/bash
# round 2, 4-GPU call: the driver's scaling invocation at N=4 (completes the 1/2/4/8 table)
O=gpurun_out/r2ad
mkdir -p $O
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 4 > $O/bench_n4.json 2> $O/bench_n4.err ) 2> $O/bench_n4_time.txt
tail -3 $O/bench_n4_time.txt
python - <<PY
import json
try:
    txt=[l for l in open("$O/bench_n4.json") if l.startswith("{")][-1]
    d=json.loads(txt)
    print(d["dtype"], d["n_gpus"], {k:round(d[k],3) for k in ("value","ms_per_step")}, "e2e", round(d["e2e"]["value"]), "ranks", [round(x,3) for x in d["rank_ms_per_step"]], "bf16", round(d["bf16"]["value"]), round(d["bf16"]["ms_per_step"],3))
    print("train", {k:(round(v["ms_step"],2), round(v["img_s"]), round(v["allreduce_exposed_ms"],2)) for k,v in d["train"].items() if isinstance(v,dict)})
except Exception as e:
    print("ERR", e); print(open("$O/bench_n4.err").read()[-3000:])
PY
