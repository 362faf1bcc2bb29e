#!/bin/bash
FZ_TEST_TIMEOUT=240 bash scripts/gpu_tests.sh 2>&1 | grep -E "exit|passed|failed|Error|error" | head -20
timeout -k 5 120 python scripts/probe_stats.py maxsim
timeout -k 5 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --systems dpr,colbert > gpurun_out/b_dc2.json 2> gpurun_out/b_dc2.err; tail -2 gpurun_out/b_dc2.err; python scripts/bench_summary.py gpurun_out/b_dc2.json
