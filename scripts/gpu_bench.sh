#!/bin/bash
# One bench.py run + a one-screen summary: bash scripts/gpu_bench.sh <tag> [bench.py args ...]
mkdir -p gpurun_out
tag=$1; shift
timeout -k 5 ${FZ_BENCH_TIMEOUT:-1500} python bench.py "$@" > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
echo "rc $?"; tail -4 gpurun_out/bench_$tag.err
python - gpurun_out/bench_$tag.json <<'PY'
import json, sys
d = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
print("QPS", round(d["value"]), "ms/step", round(d["ms_per_step"], 1), "e2e QPS", round(d["e2e"]["value"]), "e2e ms", round(d["e2e"]["ms_per_step"], 1),
      "h2d MB", round(d["e2e"]["h2d_bytes_per_step"] / 1e6, 1))
print("stage", {k: round(v, 1) for k, v in d["stage_ms"].items()})
print("kern", {k: round(v["ms"], 1) for k, v in sorted(d["kernel_ms"].items(), key=lambda x: -x[1]["ms"])})
print("frac", {k: round(v["frac"], 3) for k, v in d["kernels"].items()})
print("per_system_qps", {k: round(v) for k, v in d.get("per_system_qps", {}).items()})
print("parity", d.get("parity"))
print("build_s", {k: round(v, 1) for k, v in d["config"]["index_build_s"].items()}, "gen_s", {k: round(v, 1) for k, v in d["config"]["synth_generation_s"].items()},
      "colbert docs", d["config"]["colbert_store_docs"])
print("clocks", d["clocks"])
PY
