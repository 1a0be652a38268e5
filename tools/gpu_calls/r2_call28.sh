#!/bin/bash
# round 2, call 28: training step after the reduce change (tests + kernel-time breakdown)
O=gpurun_out/r2ab
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_backward.py tests/test_gpu_loss.py -m gpu -q > $O/pytest_bwd.log 2>&1; echo "pytest exit $?" >> $O/pytest_bwd.log
timeout 600 python tools/train_profile.py > $O/train_profile_bf16.txt 2>&1
tail -3 $O/pytest_bwd.log | cut -c1-200; head -30 $O/train_profile_bf16.txt | cut -c1-150
