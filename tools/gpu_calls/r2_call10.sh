#!/bin/bash
# round 2, call 10: chain kernel -- first run (descriptor semantics), parity, micro-benchmark
O=gpurun_out/r2j
mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_chain.py -m gpu -x -q > $O/pytest_chain.log 2>&1; echo "pytest exit $?" >> $O/pytest_chain.log
DYNMM_CHAIN_DESC_SWAP=1 timeout 300 python -m pytest tests/test_gpu_chain.py -m gpu -x -q > $O/pytest_chain_swap.log 2>&1; echo "pytest exit $?" >> $O/pytest_chain_swap.log
timeout 300 python tools/chain_bench.py > $O/chain_bench.txt 2>&1
tail -n 15 $O/pytest_chain.log | cut -c1-300
echo ---- swap; tail -n 6 $O/pytest_chain_swap.log | cut -c1-300
echo ---- bench; cat $O/chain_bench.txt | tail -12
