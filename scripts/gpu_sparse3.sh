#!/bin/bash
mkdir -p gpurun_out
FZ_TEST_TIMEOUT=200 bash scripts/gpu_tests.sh 2>&1 | grep -E "exit|passed|failed|Error|error|assert" | head -30
B="timeout -k 5 200 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --docs 2000000 --queries 2048"
for cfg in "2048 2048" "2048 4096" "1024 4096"; do
  set -- $cfg
  FZ_TILE_DOCS_LEX=$1 FZ_TILE_DOCS_SP=$2 $B --systems bm25,splade > gpurun_out/b_s3.json 2> gpurun_out/b_s3.err; echo "== lex tile $1 sp tile $2 rc $?"; python scripts/bench_summary.py gpurun_out/b_s3.json | grep -E "kernel_ms"
done
B="timeout -k 5 280 python bench.py --steps 2 --warmup 1 --no-cpu-baseline"
FZ_TILE_DOCS_SP=4096 $B --systems bm25,splade > gpurun_out/b_ls.json 2> gpurun_out/b_ls.err; echo "== full bm25+splade rc $?"; tail -3 gpurun_out/b_ls.err; python scripts/bench_summary.py gpurun_out/b_ls.json
