#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --docs 2000000 --queries 2048 --systems bm25,splade"
timeout 200 $CMD > gpurun_out/plain_sp.json 2> gpurun_out/plain_sp.err || { echo "plain run failed"; tail -5 gpurun_out/plain_sp.err; exit 1; }
NCU="ncu --set full --clock-control none --import-source on"
timeout 250 $NCU -k regex:sparse_tile_kernel -s 4 -c 1 -f -o gpurun_out/prof4_sparse_f64 $CMD > gpurun_out/ncu4_sparse_f64.log 2>&1; echo f64 $?
timeout 250 $NCU -k regex:sparse_tile_kernel -s 9 -c 1 -f -o gpurun_out/prof4_sparse_f32 $CMD > gpurun_out/ncu4_sparse_f32.log 2>&1; echo f32 $?
