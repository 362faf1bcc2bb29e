import json, sys
d = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
print("QPS", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 1), "e2e QPS", round(d["e2e"]["value"], 1))
print("stage_ms", {k: round(v, 1) for k, v in d["stage_ms"].items()})
print("kernel_ms", {k: (v["launches"], round(v["ms"], 2)) for k, v in d["kernel_ms"].items()})
print("roofline frac", {k: round(v["frac"], 3) for k, v in d["kernels"].items()})
print("clocks", d["clocks"])
