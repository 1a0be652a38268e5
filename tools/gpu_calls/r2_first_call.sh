#!/bin/bash
# First GPU call of round 2 (one B200, ~6 min): everything written after round 1's GPU budget ran out gets its first
# run, plus the standing numbers.  Outputs in gpurun_out/r2a/.
O=gpurun_out/r2a
mkdir -p $O
timeout 500 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.log
timeout 300 python -m pytest tests -m "gpu and first_run" -q -rA > $O/pytest_first_run.log 2>&1; echo "pytest exit $?" >> $O/pytest_first_run.log
timeout 300 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err
timeout 300 python bench.py --batch 32 --steps 60 --no-cpu-baseline > $O/bench_b32.json 2> $O/bench_b32.err
DYNMM_CONV_WIDE=1 timeout 200 python bench.py --no-cpu-baseline --steps 100 > $O/bench_conv_wide.json 2> $O/bench_conv_wide.err
timeout 60 python tools/pair_bench.py > $O/pair_bench.txt 2>&1
DYNMM_PAIR_ROT=1 timeout 60 python tools/pair_bench.py > $O/pair_bench_rot.txt 2>&1
DYNMM_PAIR_ROT=1 timeout 200 python bench.py --no-cpu-baseline --steps 100 > $O/bench_pair_rot.json 2> $O/bench_pair_rot.err
timeout 120 python tools/ce_bench.py > $O/ce_bench.txt 2>&1
timeout 200 python tools/modality_bench.py > $O/modality_bench.txt 2>&1
timeout 200 python tools/noise_sweep.py --batches 24 > $O/noise_sweep.txt 2>&1
timeout 200 python tools/noise_sweep.py --batches 24 --labels > $O/noise_sweep_labels.txt 2>&1
timeout 300 ncu --profile-from-start off --clock-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --csv --log-file $O/step_metrics.csv python tools/one_step.py > $O/one_step.log 2>&1
timeout 200 python tools/train_profile.py > $O/train_profile_bf16.txt 2>&1
TRAIN_PRECISION=bf16 timeout 150 python tools/train_step_dp.py > $O/train_bf16.log 2>&1
PER_GPU_BATCH=32 TRAIN_PRECISION=bf16 timeout 200 python tools/train_step_dp.py > $O/train_bf16_b32.log 2>&1
timeout 120 python tools/conv_shapes.py > $O/conv_shapes.txt 2>&1
tail -n 3 $O/pytest_gpu.log $O/pytest_first_run.log
cat $O/bench_n1.json | cut -c1-400
