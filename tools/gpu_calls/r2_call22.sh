#!/bin/bash
# round 2, call 22: cycle traces of split-mode (f32x3) convolutions
O=gpurun_out/r2v
mkdir -p $O
SPLIT=1 ONLY="s1 " timeout 200 python tools/conv_trace.py > $O/trace_split_s1.txt 2>&1
SPLIT=1 ONLY="s3 1x3 c256" timeout 200 python tools/conv_trace.py > $O/trace_split_s3.txt 2>&1
ONLY="s1 3x1" timeout 200 python tools/conv_trace.py > $O/trace_bf16_s1.txt 2>&1
cat $O/trace_split_s1.txt | head -52; head -18 $O/trace_split_s3.txt; head -18 $O/trace_bf16_s1.txt
