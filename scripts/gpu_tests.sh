#!/bin/bash
# Run the GPU parity tests, one pytest process per file so a faulting kernel cannot take the others down.
# Usage (on the GPU box):  bash scripts/gpu_tests.sh [extra pytest args]
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
rc=0
for f in tests/test_gpu_select_fuse.py tests/test_gpu_lexical.py tests/test_gpu_dense_maxsim.py; do
  name=$(basename $f .py)
  timeout 900 python -m pytest $f -m gpu -q --timeout 240 -p no:cacheprovider "$@" > gpurun_out/$name.log 2>&1
  r=$?
  echo "== $f exit $r"; tail -n 25 gpurun_out/$name.log
  [ $r -ne 0 ] && rc=1
done
exit $rc
