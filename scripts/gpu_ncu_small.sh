#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --ncu-range --systems dpr,splade"
timeout 300 $CMD > gpurun_out/plain_small.json 2> gpurun_out/plain_small.err || { echo "plain run failed"; tail -5 gpurun_out/plain_small.err; exit 1; }
NCU="ncu --set full --clock-control none --import-source on --profile-from-start off"
timeout 400 $NCU -k regex:cand_select_kernel -s 5 -c 1 -f -o gpurun_out/prof_candsel $CMD > gpurun_out/ncu_candsel.log 2>&1; echo candsel $?
timeout 400 $NCU -k regex:fuse_kernel -c 1 -f -o gpurun_out/prof_fuse $CMD > gpurun_out/ncu_fuse.log 2>&1; echo fuse $?
