#!/bin/bash
# what the driver runs at round end: smoke(), the GPU tests, the default benchmark (both arms)
mkdir -p gpurun_out
timeout -k 5 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
FZ_TEST_TIMEOUT=240 bash scripts/gpu_tests.sh 2>&1 | grep -E "exit|passed|failed|Error|error|assert" | head -30
timeout -k 5 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc $?"; python scripts/bench_summary.py gpurun_out/bench_default.json
timeout -k 5 900 python bench.py --impl reference > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "reference rc $?"; head -c 1500 gpurun_out/bench_reference.json
