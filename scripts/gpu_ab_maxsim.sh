#!/bin/bash
# A/B of the MaxSim ring depth on the stand-alone probe (scripts/probe_stats.py maxsim)
for s in "$@"; do
    echo "== FZ_MS_STAGES=$s"
    FZ_MS_STAGES=$s timeout -k 5 120 python scripts/probe_stats.py maxsim 2>&1 | grep -E "^maxsim|Error|error" | head -3
done
