#!/bin/bash
# round 2, call 11: chain kernel cycle traces, exchange cost
O=gpurun_out/r2k
mkdir -p $O
timeout 300 python tools/chain_trace.py > $O/chain_trace.txt 2>&1
DYNMM_CHAIN_NOEXCH=1 timeout 300 python tools/chain_bench.py > $O/chain_bench_noexch.txt 2>&1
cat $O/chain_trace.txt | head -120
echo ---- noexch; cat $O/chain_bench_noexch.txt | tail -8
