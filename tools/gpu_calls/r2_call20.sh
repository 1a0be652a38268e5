#!/bin/bash
# round 2, call 20: ncu evidence of the final state (both engine precisions), launch list of the bench command
O=gpurun_out/r2t
mkdir -p $O
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active
PRECISION=f32x3 timeout 400 ncu --profile-from-start off --clock-control none --metrics $M --csv --log-file $O/r2_f32x3_step_metrics.csv python tools/one_step.py > $O/one_step_f32x3.log 2>&1
PRECISION=bf16 timeout 400 ncu --profile-from-start off --clock-control none --metrics $M --csv --log-file $O/r2_bf16_step_metrics.csv python tools/one_step.py > $O/one_step_bf16.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r2_bench_launches.csv python bench.py --steps 2 --warmup 3 --min-seconds 0 --no-train --no-eager --no-cpu-baseline > $O/bench_under_ncu.log 2>&1
PRECISION=f32x3 timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:conv_igemm -s 60 -c 1 -o $O/r2_f32x3_conv_full python tools/one_step.py > $O/ncu_full.log 2>&1
PRECISION=bf16 timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:conv_chain -c 1 -o $O/r2_bf16_chain_full python tools/one_step.py > $O/ncu_full_chain.log 2>&1
tail -2 $O/one_step_f32x3.log; tail -2 $O/one_step_bf16.log; tail -3 $O/bench_under_ncu.log | cut -c1-300; tail -2 $O/ncu_full.log; tail -2 $O/ncu_full_chain.log; ls -la $O
