#!/bin/bash
# `--set full` capture (with source-level stall sampling) of the largest SPLADE head-GEMM launch of one step
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --parity-queries 0 --ncu-range --systems splade"
NCU="ncu --set full --clock-control none --import-source on --profile-from-start off"
timeout 600 $NCU -k regex:filter_gemm_kernel -s 4 -c 1 -f -o gpurun_out/prof_r02d_head_gemm $CMD > gpurun_out/ncu_head.log 2>&1; echo head $?
ls -la gpurun_out/prof_r02d_*.ncu-rep
