#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --systems dpr,colbert --docs 1000000"
for pool in 2000 20000 200000 1000000; do
  $B --pool $pool > gpurun_out/b_pool$pool.json 2> gpurun_out/b_pool$pool.err; echo pool $pool; python scripts/bench_summary.py gpurun_out/b_pool$pool.json | grep -E "kernel_ms|roofline"
done
