#!/bin/bash
# Layout-knob sweep of the two sparse kernels at full size (BM25 + SPLADE only); prints the per-kernel ms of each setting.
mkdir -p gpurun_out
run() {
    tag=$1; shift
    env "$@" timeout -k 5 300 python bench.py --systems bm25,splade --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/sweep_$tag.json 2> gpurun_out/sweep_$tag.err
    python - "$tag" gpurun_out/sweep_$tag.json <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[2]) if l.startswith("{")][-1])
    k = d["kernel_ms"]
    print(sys.argv[1], "ms/step", round(d["ms_per_step"], 1), {n: round(k[n]["ms"], 1) for n in k if n.startswith("sparse_tile") or n.startswith("cand")})
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
}
for cfg in "$@"; do
    tag=$(echo "$cfg" | tr '= ,/' '____')
    run "$tag" $(echo "$cfg" | tr ',' ' ')
done
