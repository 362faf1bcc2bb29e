#!/bin/bash
# final 1-GPU pass: smoke, all GPU tests, default bench (both arms), then the launch list and the head-GEMM capture of the final build
bash scripts/gpu_final.sh
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --parity-queries 0 --ncu-range"
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1; echo launches $?
NCU="ncu --set full --clock-control none --import-source on --profile-from-start off"
timeout 600 $NCU -k regex:filter_gemm_kernel -s 12 -c 1 -f -o gpurun_out/prof_r02f_head_gemm $CMD > gpurun_out/ncu_head.log 2>&1; echo head $?
