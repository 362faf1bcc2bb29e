#!/bin/bash
# Run the GPU parity tests, one pytest process per file so a faulting kernel cannot take the others down.
# Every process runs under a hard `timeout`: a hung kernel must not eat the GPU budget.
# Usage (on the GPU box):  bash scripts/gpu_tests.sh [extra pytest args]
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
rc=0
for f in tests/test_gpu_*.py; do
  name=$(basename $f .py)
  timeout -k 5 ${FZ_TEST_TIMEOUT:-240} python -u -m pytest $f -m gpu -q --timeout 120 -p no:cacheprovider "$@" > gpurun_out/$name.log 2>&1
  r=$?
  echo "== $f exit $r"; tail -n 25 gpurun_out/$name.log
  [ $r -ne 0 ] && rc=1
done
exit $rc
