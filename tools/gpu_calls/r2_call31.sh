#!/bin/bash
# round 2, call 31: ppm-1-2-4-8 on the engine + regression (whole suite twice: flakiness check of the flag-based kernels)
O=gpurun_out/r2ae
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu_1.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu_1.log
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu_2.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu_2.log
for i in 1 2 3 4 5 6; do timeout 300 python -m pytest tests/test_gpu_chain.py -m gpu -q > $O/chain_$i.log 2>&1; tail -1 $O/chain_$i.log; done
tail -n 3 $O/pytest_gpu_1.log | cut -c1-250; tail -n 3 $O/pytest_gpu_2.log | cut -c1-250
