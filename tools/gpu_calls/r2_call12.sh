#!/bin/bash
# round 2, call 12: chain kernel with pipelined epilogue + asynchronous halo exchange
O=gpurun_out/r2l
mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_chain.py -m gpu -q > $O/pytest_chain.log 2>&1; echo "pytest exit $?" >> $O/pytest_chain.log
timeout 300 python tools/chain_bench.py > $O/chain_bench.txt 2>&1
timeout 300 python tools/chain_trace.py > $O/chain_trace.txt 2>&1
tail -n 12 $O/pytest_chain.log | cut -c1-300
echo ---- bench; cat $O/chain_bench.txt | tail -8
grep -A 13 "stage3 rgb+depth(4)\|decoder 30x40\|stage2" $O/chain_trace.txt
