#!/bin/bash
# round 2, call 16: engine-backed local gate (first run) + the whole GPU suite
O=gpurun_out/r2p
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_local_gate.py -m gpu -q > $O/pytest_local.log 2>&1; echo "pytest exit $?" >> $O/pytest_local.log
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.log
tail -n 40 $O/pytest_local.log | cut -c1-250
tail -n 6 $O/pytest_gpu.log | cut -c1-250
