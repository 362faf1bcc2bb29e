#!/bin/bash
# ncu captures of the four scoring kernels on the full-size benchmark (one GPU).  Plain run first, as required.
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
$CMD > gpurun_out/plain.json 2> gpurun_out/plain.err || { echo "plain run failed"; tail -5 gpurun_out/plain.err; exit 1; }
python scripts/bench_summary.py gpurun_out/plain.json
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:sparse_tile_kernel -s 6 -c 1 -f -o gpurun_out/prof_sparse_f64 $CMD > gpurun_out/ncu_sparse_f64.log 2>&1; echo f64 $?
$NCU -k regex:sparse_tile_kernel -s 13 -c 1 -f -o gpurun_out/prof_sparse_f32 $CMD > gpurun_out/ncu_sparse_f32.log 2>&1; echo f32 $?
$NCU -k regex:dense_filter_kernel -s 7 -c 1 -f -o gpurun_out/prof_dense $CMD > gpurun_out/ncu_dense.log 2>&1; echo dense $?
$NCU -k regex:maxsim_kernel -c 1 -f -o gpurun_out/prof_maxsim $CMD > gpurun_out/ncu_maxsim.log 2>&1; echo maxsim $?
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1; echo launches $?
ls -la gpurun_out/*.ncu-rep
