#!/bin/bash
# round 2, call 37: local-gate engine throughput at the bench shape; re-check of the last test edits
O=gpurun_out/r2ak
mkdir -p $O
timeout 600 python tools/local_gate_bench.py > $O/local_gate_bench.txt 2> $O/local_gate_bench.err
timeout 900 python -m pytest tests/test_gpu_f32x3.py tests/test_gpu_local_gate.py tests/test_gpu_loss.py tests/test_gpu_pair.py tests/test_gpu_robustness.py -m gpu -q > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log
tail -2 $O/local_gate_bench.txt | cut -c1-600; tail -3 $O/local_gate_bench.err | cut -c1-300; grep -E "passed|failed|FAILED" $O/pytest.log | tail -3
