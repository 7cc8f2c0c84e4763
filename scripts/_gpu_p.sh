#!/bin/bash
python -m pytest tests -x -q -m gpu -k "c5 or compact or scattered or host or stream or lazy" > gpurun_out/r02p_tests.log 2>&1; tail -3 gpurun_out/r02p_tests.log
{
for wl in C5dis C5nn; do
  bash scripts/ab.sh "$wl f32" --workload $wl --steps 20 --warmup 3
  bash scripts/ab.sh "$wl f64 data, compact" --workload $wl --kernel compact --xdtype f64 --steps 10 --warmup 3
done
} > gpurun_out/r02p_ab_compact_pass2_occupancy.txt 2>&1
cat gpurun_out/r02p_ab_compact_pass2_occupancy.txt
python scripts/call_latency.py > gpurun_out/r02p_call_latency.txt 2>/dev/null; cat gpurun_out/r02p_call_latency.txt
