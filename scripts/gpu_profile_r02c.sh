#!/bin/bash
# Round-2 follow-up profile (after the DPR epilogue rewrite): plain full-size run, ncu launch list of one timed step, and one
# `--set full` capture of the largest DPR filter-GEMM launch (7th launch of filter_gemm_kernel: the 3.98M-doc round).
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --parity-queries 0 --ncu-range"
timeout 400 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --parity-queries 0 > gpurun_out/plain.json 2> gpurun_out/plain.err || { echo "plain run failed"; tail -5 gpurun_out/plain.err; exit 1; }
python scripts/bench_summary.py gpurun_out/plain.json
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1; echo launches $?
NCU="ncu --set full --clock-control none --import-source on --profile-from-start off"
timeout 600 $NCU -k regex:filter_gemm_kernel -s 6 -c 1 -f -o gpurun_out/prof_r02b_dense $CMD > gpurun_out/ncu_dense.log 2>&1; echo dense $?
ls -la gpurun_out/prof_r02b_*.ncu-rep
