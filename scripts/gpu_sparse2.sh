#!/bin/bash
mkdir -p gpurun_out
bash scripts/gpu_tests.sh 2>&1 | grep -E "exit|passed|failed|Error|error|assert" | head -30
B="timeout -k 5 280 python bench.py --steps 2 --warmup 1 --no-cpu-baseline"
$B --systems bm25 > gpurun_out/b_lex.json 2> gpurun_out/b_lex.err; echo "== bm25 rc $?"; tail -3 gpurun_out/b_lex.err; python scripts/bench_summary.py gpurun_out/b_lex.json
$B --systems splade > gpurun_out/b_sp.json 2> gpurun_out/b_sp.err; echo "== splade rc $?"; tail -3 gpurun_out/b_sp.err; python scripts/bench_summary.py gpurun_out/b_sp.json
