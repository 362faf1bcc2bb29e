#!/bin/bash
mkdir -p gpurun_out
FZ_TEST_TIMEOUT=200 bash scripts/gpu_tests.sh 2>&1 | grep -E "exit|passed|failed|Error|error|assert" | head -30
B="timeout -k 5 280 python bench.py --steps 2 --warmup 1 --no-cpu-baseline"
for t in 2048 4096; do
FZ_TILE_DOCS_SP=$t $B --systems splade > gpurun_out/b_sp$t.json 2> gpurun_out/b_sp$t.err; echo "== full splade tile $t rc $?"; tail -3 gpurun_out/b_sp$t.err; python scripts/bench_summary.py gpurun_out/b_sp$t.json | grep -E "kernel_ms|roofline"
done
for t in 1024 2048; do
FZ_TILE_DOCS_LEX=$t $B --systems bm25 > gpurun_out/b_lex$t.json 2> gpurun_out/b_lex$t.err; echo "== full bm25 tile $t rc $?"; tail -3 gpurun_out/b_lex$t.err; python scripts/bench_summary.py gpurun_out/b_lex$t.json | grep -E "kernel_ms|roofline"
done
