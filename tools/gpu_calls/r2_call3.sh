#!/bin/bash
# round 2, call 3: tile-completion flags (layer overlap), in-kernel traces of the streamed-weight layers
O=gpurun_out/r2c
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.log
for d in 0 1; do
  DUAL=$d ONLY="s3 1x3 c256" timeout 120 python tools/conv_trace.py > $O/trace_s3_dual$d.txt 2>&1
  DUAL=$d ONLY="s4 1x3 c512" timeout 120 python tools/conv_trace.py > $O/trace_s4_dual$d.txt 2>&1
done
ONLY="s2 1x3 c128" timeout 120 python tools/conv_trace.py > $O/trace_s2.txt 2>&1
DYNMM_TILE_FLAGS=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-train --no-eager --no-cpu-baseline > $O/bench_flags.json 2> $O/bench_flags.err
DYNMM_TILE_FLAGS=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-train --no-eager --no-cpu-baseline > $O/bench_noflags.json 2> $O/bench_noflags.err
DYNMM_TILE_FLAGS=1 DYNMM_CONV_DUAL=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-train --no-eager --no-cpu-baseline > $O/bench_flags_nodual.json 2> $O/bench_flags_nodual.err
tail -n 15 $O/pytest_gpu.log
cat $O/trace_s3_dual0.txt $O/trace_s3_dual1.txt
for f in $O/bench_flags.json $O/bench_noflags.json $O/bench_flags_nodual.json; do echo $f; cut -c1-330 $f; done
tail -c 1500 $O/bench_flags.err
