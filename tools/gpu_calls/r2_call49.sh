#!/bin/bash
# round 2, call 49: stem epilogue with pairwise row exchange + lane-paired stores: stem tests and variants
O=gpurun_out/r2ao
mkdir -p $O
timeout 600 python -m pytest tests -m gpu -q -x -k "stem or golden or smoke or gate" > $O/pytest_stem.log 2>&1; echo "pytest exit $?" >> $O/pytest_stem.log
grep -E "passed|failed|FAILED" $O/pytest_stem.log | tail -4 | cut -c1-250
cp dynmm_b200/libdynmm_b200.so /tmp/new.so
for n in 0 1; do
  if [ $n = 0 ]; then cp /tmp/new.so dynmm_b200/libdynmm_b200.so; else cp tools/bin/libdynmm_stemdbg_$n.so dynmm_b200/libdynmm_b200.so; fi
  echo "variant $n: $(timeout 300 python tools/stem_bench.py 2>&1 | grep -E 'stem_s2d|r32|d16' | tr '\n' ' ')" | tee -a $O/stem_variants.txt
done
cp /tmp/new.so dynmm_b200/libdynmm_b200.so
