#!/bin/bash
# round 2, call 51: stem epilogue with 8 shuffles per map and the quad-major tile: stem tests, timing, trace
O=gpurun_out/r2aq
mkdir -p $O
timeout 600 python -m pytest tests -m gpu -q -x -k "stem or golden or smoke or gate" > $O/pytest_stem.log 2>&1; echo "pytest exit $?" >> $O/pytest_stem.log
grep -E "passed|failed|FAILED" $O/pytest_stem.log | tail -4 | cut -c1-250
timeout 300 python tools/stem_bench.py 2>&1 | tail -5 | tee $O/stem_bench.txt
cp dynmm_b200/libdynmm_b200.so /tmp/new.so
cp tools/bin/libdynmm_stemtrace.so dynmm_b200/libdynmm_b200.so
timeout 300 python tools/stem_trace.py > $O/stem_trace.txt 2>&1
cp /tmp/new.so dynmm_b200/libdynmm_b200.so
