#!/bin/bash
# round 2, call 38: different RGB / depth encoders on the engine; whole suite
O=gpurun_out/r2al
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.log
grep -E "passed|failed|^E  |FAILED" $O/pytest_gpu.log | tail -12 | cut -c1-250
