#!/bin/bash
# maxsim v2 parity + timing, f64 tile sweep
mkdir -p gpurun_out
bash scripts/gpu_tests.sh || echo "TESTS FAILED"
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline"
$B --systems dpr,colbert > gpurun_out/b_dc.json 2> gpurun_out/b_dc.err; python scripts/bench_summary.py gpurun_out/b_dc.json
for t in 1024 1536 2048; do
  FZ_TILE_DOCS_LEX=$t $B --systems bm25 > gpurun_out/b_lex$t.json 2> gpurun_out/b_lex$t.err; echo lex $t; python scripts/bench_summary.py gpurun_out/b_lex$t.json
done
for t in 3072 4096; do
  FZ_TILE_DOCS_SP=$t $B --systems splade > gpurun_out/b_sp$t.json 2> gpurun_out/b_sp$t.err; echo sp $t; python scripts/bench_summary.py gpurun_out/b_sp$t.json
done
