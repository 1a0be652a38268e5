#!/bin/bash
# round 2, call 50: clock-stamp trace of the stem epilogue (trace build in tools/bin)
O=gpurun_out/r2ap
mkdir -p $O
cp dynmm_b200/libdynmm_b200.so /tmp/new.so
cp tools/bin/libdynmm_stemtrace.so dynmm_b200/libdynmm_b200.so
timeout 300 python tools/stem_trace.py > $O/stem_trace.txt 2>&1
cp /tmp/new.so dynmm_b200/libdynmm_b200.so
tail -5 $O/stem_trace.txt
