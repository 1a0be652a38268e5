#!/bin/bash
# round 2, call 47: where does the stem epilogue spend its time?  (compile-time experiment variants in tools/bin)
O=gpurun_out/r2am
mkdir -p $O
cp dynmm_b200/libdynmm_b200.so /tmp/new.so
for n in 0 1 2 6 8; do
  if [ $n = 0 ]; then cp /tmp/new.so dynmm_b200/libdynmm_b200.so; else cp tools/bin/libdynmm_stemdbg_$n.so dynmm_b200/libdynmm_b200.so; fi
  echo "variant $n: $(timeout 300 python tools/stem_bench.py 2>&1 | grep stem_s2d)" | tee -a $O/stem_variants.txt
done
cp /tmp/new.so dynmm_b200/libdynmm_b200.so
