#!/bin/bash
# round 2, 2-GPU call: the driver's torchrun invocation of bench.py (default protocol), reference arm under torchrun
O=gpurun_out/r2u
mkdir -p $O
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > $O/bench_n2.json 2> $O/bench_n2.err ) 2> $O/bench_n2_time.txt
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > $O/bench_ref_n2.json 2> $O/bench_ref_n2.err
tail -3 $O/bench_n2_time.txt
python - <<PY
import json
try:
    txt=[l for l in open("$O/bench_n2.json") if l.startswith("{")][-1]
    d=json.loads(txt)
    print(d["dtype"], d["n_gpus"], {k:round(d[k],3) for k in ("value","ms_per_step")}, "e2e", round(d["e2e"]["value"]), "ranks", [round(x,3) for x in d["rank_ms_per_step"]], "bf16", round(d["bf16"]["value"]), round(d["bf16"]["ms_per_step"],3))
    print("train", json.dumps(d["train"])[:900])
except Exception as e:
    print("ERR", e); print(open("$O/bench_n2.err").read()[-3000:])
print(open("$O/bench_ref_n2.json").read()[:600])
PY
