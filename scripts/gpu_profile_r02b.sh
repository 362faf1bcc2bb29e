#!/bin/bash
# the three captures of scripts/gpu_profile_r02.sh whose kernel-name filter did not match (template arguments print as <0> / <1>)
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --parity-queries 0 --ncu-range"
NCU="ncu --set full --clock-control none --import-source on --profile-from-start off"
timeout 600 $NCU -k regex:filter_gemm_kernel -s 12 -c 1 -f -o gpurun_out/prof_r02_head_gemm $CMD > gpurun_out/ncu_head.log 2>&1; echo head $?
timeout 600 $NCU -k "regex:^sparse_tile_kernel$" -s 5 -c 1 -f -o gpurun_out/prof_r02_sparse_f64 $CMD > gpurun_out/ncu_f64.log 2>&1; echo f64 $?
timeout 600 $NCU -k regex:filter_gemm_kernel -s 6 -c 1 -f -o gpurun_out/prof_r02_dense $CMD > gpurun_out/ncu_dense.log 2>&1; echo dense $?
ls -la gpurun_out/prof_r02_*.ncu-rep
