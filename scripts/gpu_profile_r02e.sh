#!/bin/bash
# warp-state captures (stall reasons) of the BM25 and MaxSim kernels: is instruction fetch (stall_no_inst) an issue there too?
mkdir -p gpurun_out
NCU="ncu --section WarpStateStats --section SourceCounters --section SpeedOfLight --clock-control none --import-source on --profile-from-start off"
timeout 600 $NCU -k "regex:^sparse_tile_kernel$" -s 5 -c 1 -f -o gpurun_out/prof_r02e_bm25 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --parity-queries 0 --ncu-range --systems bm25 > gpurun_out/ncu_bm25.log 2>&1; echo bm25 $?
timeout 600 $NCU -k regex:maxsim_kernel -c 1 -f -o gpurun_out/prof_r02e_maxsim python bench.py --steps 1 --warmup 1 --no-cpu-baseline --parity-queries 0 --ncu-range --systems dpr,colbert > gpurun_out/ncu_maxsim.log 2>&1; echo maxsim $?
timeout 600 $NCU -k regex:tail_codes_kernel -s 4 -c 1 -f -o gpurun_out/prof_r02e_tail python bench.py --steps 1 --warmup 1 --no-cpu-baseline --parity-queries 0 --ncu-range --systems splade > gpurun_out/ncu_tail.log 2>&1; echo tail $?
ls -la gpurun_out/prof_r02e_*.ncu-rep
