#!/bin/bash
# round 2, call 6 (2 GPUs): the full bench line under torchrun (eval replicas + data-parallel training leg with the
# hook-driven bucketed all-reduce), after a single-GPU regression check of the restructured conv kernel
O=gpurun_out/r2f
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_fusion.py -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.log
CUDA_VISIBLE_DEVICES=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-train --no-eager --no-cpu-baseline > $O/bench_n1_quick.json 2> $O/bench_n1_quick.err
CUDA_VISIBLE_DEVICES=0 DYNMM_MERGE=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-train --no-eager --no-cpu-baseline > $O/bench_n1_merge.json 2> $O/bench_n1_merge.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > $O/bench_n2.json 2> $O/bench_n2.err; echo "bench n2 exit $?" >> $O/bench_n2.err
TRAIN_PRECISION=bf16 PER_GPU_BATCH=16 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/train_step_dp.py > $O/train_dp2.log 2>&1
tail -n 4 $O/pytest_gpu.log | cut -c1-200
for f in n1_quick n1_merge; do python - <<PY
import json
try:
    d=json.load(open("$O/bench_$f.json"))
    print("$f", {k:d[k] for k in ("value","ms_per_step","gpu_launches_per_step")}, {k:d["roofline"][k] for k in ("frac","kernel_s_per_step","launches_per_step")})
except Exception as e:
    print("no json", e)
PY
done
tail -c 1500 $O/bench_n2.err
cut -c1-3000 $O/bench_n2.json
tail -n 3 $O/train_dp2.log
