#!/bin/bash
# round 2, call 59: fp32-grade mode with SE-add fusion -- f32x3 and fusion test files, then the whole suite
O=gpurun_out/r2ax
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_f32x3.py -m gpu -q -x -s -k "se_" > $O/pytest_se.log 2>&1; echo "pytest exit $?" >> $O/pytest_se.log
grep -E "passed|failed|FAILED|Error|f32x3 vs" $O/pytest_se.log | tail -8 | cut -c1-300
timeout 1500 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.log
grep -E "passed|failed|FAILED|Error" $O/pytest_gpu.log | tail -4 | cut -c1-300
