#!/bin/bash
# Short round-end capture: tests, smoke, bench (ours + reference arm), ncu metrics of one step.
O=gpurun_out/final2
mkdir -p $O
timeout 400 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.log
timeout 120 python __graft_entry__.py smoke > $O/smoke.log 2>&1
timeout 300 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err
timeout 300 ncu --profile-from-start off --clock-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --csv --log-file $O/step_metrics.csv python tools/one_step.py > $O/one_step.log 2>&1
tail -n 2 $O/pytest_gpu.log; tail -n 1 $O/smoke.log
