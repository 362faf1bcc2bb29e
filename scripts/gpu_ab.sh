#!/bin/bash
# A/B of environment knobs on a subset of the systems at full size: bash scripts/gpu_ab.sh <systems> CFG [CFG ...]
# (CFG = comma-separated VAR=VALUE pairs, "-" = defaults); prints ms/step and every kernel's ms for each setting.
mkdir -p gpurun_out
systems=$1; shift
for cfg in "$@"; do
    tag=$(echo "${systems}_$cfg" | tr '= ,/' '____')
    envs=$(echo "$cfg" | tr ',' ' '); [ "$cfg" = "-" ] && envs=""
    env $envs timeout -k 5 ${FZ_AB_TIMEOUT:-400} python bench.py --systems $systems --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ab_$tag.json 2> gpurun_out/ab_$tag.err
    python - "$tag" gpurun_out/ab_$tag.json gpurun_out/ab_$tag.err <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[2]) if l.startswith("{")][-1])
    k = d["kernel_ms"]
    print(sys.argv[1], "ms/step", round(d["ms_per_step"], 1), "e2e", round(d["e2e"]["ms_per_step"], 1), "setup_s", round(d["config"]["setup_s"], 1))
    print("   kernels", {n: round(k[n]["ms"], 2) for n in sorted(k, key=lambda n: -k[n]["ms"])})
    print("   stages ", {n: round(v, 2) for n, v in d["stage_ms"].items()})
except Exception as e:
    print(sys.argv[1], "FAILED", e)
    print(open(sys.argv[3]).read()[-1500:])
PY
done
