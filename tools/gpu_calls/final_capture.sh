#!/bin/bash
# Round-end capture on one B200: tests, bench (ours + reference arm), ncu launch list / metrics of one step,
# full ncu capture of the stem, UMMA issue microbenchmark, convolution-program timeline, training step.
# Everything lands in gpurun_out/final/ (summaries are copied to profiles/ by hand).
O=gpurun_out/final
mkdir -p $O
timeout 400 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.log
timeout 300 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err
timeout 300 ncu --profile-from-start off --clock-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --csv --log-file $O/step_metrics.csv python tools/one_step.py > $O/one_step.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:stem_s2d_kernel -c 1 -o $O/stem_full python tools/one_step.py > $O/stem_full.log 2>&1
timeout 60 ./tools/bin/umma_issue_bench > $O/umma_issue_bench.txt 2>&1
DYNMM_PROGRAM=1 timeout 120 python tools/program_timeline.py > $O/program_timeline.txt 2>&1
DYNMM_PROGRAM=1 timeout 200 python bench.py --no-cpu-baseline > $O/bench_program_path.json 2> $O/bench_program_path.err
TRAIN_PRECISION=bf16 timeout 150 python tools/train_step_dp.py > $O/train_bf16.log 2>&1
TRAIN_PRECISION=autocast timeout 150 python tools/train_step_dp.py > $O/train_autocast.log 2>&1
tail -n 2 $O/pytest_gpu.log
tail -n 1 $O/train_bf16.log $O/train_autocast.log
