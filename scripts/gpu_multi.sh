#!/bin/bash
# 2-GPU checks: parity of the sharded pipeline against one index, then the benchmark launched the way the driver does.
mkdir -p gpurun_out
N=${1:-2}
timeout -k 5 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 scripts/multi_gpu_parity.py > gpurun_out/parity_g$N.log 2>&1; echo "parity rc $?"; grep -cE ": OK" gpurun_out/parity_g$N.log; grep -E "MISMATCH|PARITY|Error|error" gpurun_out/parity_g$N.log | head -20
timeout -k 5 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_g$N.json 2> gpurun_out/bench_g$N.err; echo "bench rc $?"; tail -3 gpurun_out/bench_g$N.err; python scripts/bench_summary.py gpurun_out/bench_g$N.json
