#!/bin/bash
# round 2, call 4: experiment -- every conv planned into <= 113 KiB so two CTAs (consecutive launches / both streams) share an SM
O=gpurun_out/r2d
mkdir -p $O
DYNMM_CONV_SMALL=1 timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_fusion.py -m gpu -x -q > $O/pytest_small.log 2>&1; echo "pytest exit $?" >> $O/pytest_small.log
DYNMM_CONV_SMALL=1 timeout 120 python tools/conv_shapes.py > $O/conv_shapes_small.txt 2>&1
timeout 120 python tools/conv_shapes.py > $O/conv_shapes_base.txt 2>&1
DYNMM_CONV_SMALL=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-train --no-eager --no-cpu-baseline > $O/bench_small.json 2> $O/bench_small.err
timeout 300 python bench.py --steps 20 --warmup 5 --no-train --no-eager --no-cpu-baseline > $O/bench_base.json 2> $O/bench_base.err
tail -n 4 $O/pytest_small.log
paste $O/conv_shapes_base.txt $O/conv_shapes_small.txt | cut -c1-75,120-200 | grep -v "dual=False\|dual=True"
for f in $O/bench_small.json $O/bench_base.json; do echo $f; cut -c150-330 $f; done
tail -c 800 $O/bench_small.err
