#!/bin/bash
# The `ncu --set full` captures of round_evidence.sh alone.  Usage: gpurun -- bash scripts/ncu_only.sh <tag> [workloads...]
tag=${1:-r02}; shift
wls=${@:-C4 C2 C3 C3tri C5dis C5nn}
mkdir -p gpurun_out
Q="--steps 3 --warmup 3 --no-cpu --no-e2e --no-others --strong-rows 0"
for wl in $wls; do
  python bench.py --workload $wl $Q > gpurun_out/${tag}_plain_${wl}.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:'staged_kernel|compact|gather_kernel' -s 3 -c $([ $wl = C5dis ] && echo 2 || echo 1) -f \
      -o gpurun_out/${tag}_${wl} python bench.py --workload $wl $Q > gpurun_out/${tag}_ncu_${wl}.log 2>&1
  tail -2 gpurun_out/${tag}_ncu_${wl}.log
done
ls -la gpurun_out/*.ncu-rep | tail
