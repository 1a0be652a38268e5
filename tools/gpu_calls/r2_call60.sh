#!/bin/bash
# round 2, call 60: fp32-grade mode for the local-gate engine -- local-gate tests, then the whole suite
O=gpurun_out/r2ay
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_local_gate.py -m gpu -q -x > $O/pytest_local.log 2>&1; echo "pytest exit $?" >> $O/pytest_local.log
grep -E "passed|failed|FAILED|Error" $O/pytest_local.log | tail -6 | cut -c1-300
timeout 1500 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.log
grep -E "passed|failed|FAILED|Error" $O/pytest_gpu.log | tail -4 | cut -c1-300
