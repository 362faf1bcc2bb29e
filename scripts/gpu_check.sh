#!/bin/bash
# parity tests + one full-size benchmark step pair (no CPU baseline)
mkdir -p gpurun_out
FZ_TEST_TIMEOUT=240 bash scripts/gpu_tests.sh 2>&1 | grep -E "exit|passed|failed|Error|error|assert" | head -30
timeout -k 5 400 python bench.py --steps 2 --warmup 1 --no-cpu-baseline "$@" > gpurun_out/check.json 2> gpurun_out/check.err; echo "bench rc $?"; tail -3 gpurun_out/check.err; python scripts/bench_summary.py gpurun_out/check.json
