#!/bin/bash
# Round-1 profile pass on ONE B200: plain full-size run first (as required), then the ncu launch list of one timed step
# (cudaProfilerStart/Stop bracket, --ncu-range) and one `--set full` capture per scoring kernel (the largest round: the step
# runs DPR, SPLADE (fp32 launches 0-6), BM25 (fp64 launches 7-13), MaxSim).
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --ncu-range"
timeout 400 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/plain.json 2> gpurun_out/plain.err || { echo "plain run failed"; tail -5 gpurun_out/plain.err; exit 1; }
python scripts/bench_summary.py gpurun_out/plain.json
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1; echo launches $?
NCU="ncu --set full --clock-control none --import-source on --profile-from-start off"
timeout 600 $NCU -k regex:sparse_tile_kernel -s 12 -c 1 -f -o gpurun_out/prof_sparse_f64 $CMD > gpurun_out/ncu_sparse_f64.log 2>&1; echo f64 $?
timeout 600 $NCU -k regex:sparse_tile_kernel -s 5 -c 1 -f -o gpurun_out/prof_sparse_f32 $CMD > gpurun_out/ncu_sparse_f32.log 2>&1; echo f32 $?
timeout 600 $NCU -k regex:dense_filter_kernel -s 6 -c 1 -f -o gpurun_out/prof_dense $CMD > gpurun_out/ncu_dense.log 2>&1; echo dense $?
timeout 600 $NCU -k regex:maxsim_kernel -c 1 -f -o gpurun_out/prof_maxsim $CMD > gpurun_out/ncu_maxsim.log 2>&1; echo maxsim $?
ls -la gpurun_out/*.ncu-rep
