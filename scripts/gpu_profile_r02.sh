#!/bin/bash
# Round-2 profile pass on ONE B200: plain full-size run first, then the ncu launch list of one timed step (cudaProfilerStart /
# Stop bracket, --ncu-range) and one `--set full` capture of the largest launch of every scoring kernel.
# Launch order of a step: DPR (filter_gemm<false> x8 rounds), SPLADE (bootstrap: sparse_tile_fx x4; then per round tail_codes,
# filter_gemm<true>, sparse_rescore), BM25 (sparse_tile_kernel<double> x6), MaxSim.
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --parity-queries 0 --ncu-range"
timeout 400 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --parity-queries 0 > gpurun_out/plain.json 2> gpurun_out/plain.err || { echo "plain run failed"; tail -5 gpurun_out/plain.err; exit 1; }
python scripts/bench_summary.py gpurun_out/plain.json
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1; echo launches $?
NCU="ncu --set full --clock-control none --import-source on --profile-from-start off"
timeout 600 $NCU -k regex:tail_codes_kernel -s 4 -c 1 -f -o gpurun_out/prof_r02_tail_codes $CMD > gpurun_out/ncu_tail.log 2>&1; echo tail $?
timeout 600 $NCU -k regex:filter_gemm_kernel -s 12 -c 1 -f -o gpurun_out/prof_r02_head_gemm $CMD > gpurun_out/ncu_head.log 2>&1; echo head $?
timeout 600 $NCU -k regex:sparse_rescore_kernel -s 5 -c 1 -f -o gpurun_out/prof_r02_rescore $CMD > gpurun_out/ncu_rescore.log 2>&1; echo rescore $?
timeout 600 $NCU -k "regex:^sparse_tile_kernel$" -s 5 -c 1 -f -o gpurun_out/prof_r02_sparse_f64 $CMD > gpurun_out/ncu_f64.log 2>&1; echo f64 $?
timeout 600 $NCU -k regex:filter_gemm_kernel -s 6 -c 1 -f -o gpurun_out/prof_r02_dense $CMD > gpurun_out/ncu_dense.log 2>&1; echo dense $?
timeout 600 $NCU -k regex:maxsim_kernel -c 1 -f -o gpurun_out/prof_r02_maxsim $CMD > gpurun_out/ncu_maxsim.log 2>&1; echo maxsim $?
ls -la gpurun_out/prof_r02_*.ncu-rep
