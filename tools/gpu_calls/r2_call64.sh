#!/bin/bash
# round 2, call 64: per-tile trace of the split (f32x3) convolution at the stage-1 and stage-2 shapes
O=gpurun_out/r2ba
mkdir -p $O
cp dynmm_b200/libdynmm_b200.so /tmp/new.so
cp tools/bin/libdynmm_tiles.so dynmm_b200/libdynmm_b200.so
SPLIT=1 TILES=1 ONLY="s1 1x3 c64" timeout 200 python tools/conv_trace.py > $O/trace_tiles_split_s1.txt 2>&1
TILES=1 ONLY="s1 1x3 c64" timeout 200 python tools/conv_trace.py > $O/trace_tiles_bf16_s1.txt 2>&1
SPLIT=1 TILES=1 ONLY="s2" timeout 200 python tools/conv_trace.py > $O/trace_tiles_split_s2.txt 2>&1
cp /tmp/new.so dynmm_b200/libdynmm_b200.so
head -45 $O/trace_tiles_split_s1.txt
