#!/bin/bash
# round 2, call 39: CUDA-graph replay of the local-gate engine
O=gpurun_out/r2am
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_local_gate.py -m gpu -q > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log
timeout 600 python tools/local_gate_bench.py > $O/local_gate_bench.txt 2> $O/local_gate_bench.err
grep -E "passed|failed|^E  |FAILED" $O/pytest.log | tail -8 | cut -c1-250; tail -1 $O/local_gate_bench.txt | cut -c1-700; tail -2 $O/local_gate_bench.err | cut -c1-300
