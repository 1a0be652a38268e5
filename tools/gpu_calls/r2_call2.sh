#!/bin/bash
# round 2, call 2: tests after the host-side changes + dual-M units, the new bench line (eager baseline, train leg),
# per-shape conv timings with and without dual-M units, the TMA ingest micro-benchmark
O=gpurun_out/r2b
mkdir -p $O
timeout 600 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.log
timeout 120 tools/bin/tma_bw_bench > $O/tma_bw_bench.txt 2>&1
timeout 120 python tools/conv_shapes.py > $O/conv_shapes.txt 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench exit $?" >> $O/bench_n1.err
DYNMM_CONV_DUAL=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-train --no-eager --no-cpu-baseline > $O/bench_nodual.json 2> $O/bench_nodual.err
DYNMM_CONV_DUAL=2 timeout 300 python bench.py --steps 20 --warmup 5 --no-train --no-eager --no-cpu-baseline > $O/bench_dual2.json 2> $O/bench_dual2.err
timeout 200 python bench.py --impl reference --steps 20 --warmup 5 > $O/bench_reference.json 2> $O/bench_reference.err
tail -n 5 $O/pytest_gpu.log
tail -c 2000 $O/bench_n1.err
cat $O/bench_n1.json | cut -c1-1500
cat $O/tma_bw_bench.txt | tail -40
