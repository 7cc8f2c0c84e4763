#!/bin/bash
# A/B helper for the GPU box: ab.sh <label> <bench args...>  -> one line "label GB/s ms launches"
label=$1; shift
out=$(python bench.py --no-cpu --no-e2e --no-others --strong-rows 0 "$@" 2>gpurun_out/ab_err.log | tail -1)
python - "$label" "$out" <<'PY'
import json, sys
try:
    d = json.loads(sys.argv[2])
    print(sys.argv[1], round(d["roofline"]["achieved"]), round(d["ms_per_step"], 4), d["gpu_launches"])
except Exception as e:
    print(sys.argv[1], "FAILED", sys.argv[2][:200], open("gpurun_out/ab_err.log").read()[-400:])
PY
