#!/bin/bash
python -m pytest tests -x -q -m gpu -k "c5 or compact or scattered or random_operator or named_configs or levels_mixed" > gpurun_out/r02q_tests.log 2>&1; tail -3 gpurun_out/r02q_tests.log
{
for wl in C5dis C5nn; do
  bash scripts/ab.sh "$wl f32" --workload $wl --steps 20 --warmup 3
  bash scripts/ab.sh "$wl f64 data, compact" --workload $wl --kernel compact --xdtype f64 --steps 10 --warmup 3
  bash scripts/ab.sh "$wl f32 batch 32" --workload $wl --kernel compact --batch 32 --steps 20 --warmup 3
done
} > gpurun_out/r02q_ab_compact_pass2_persistent.txt 2>&1
cat gpurun_out/r02q_ab_compact_pass2_persistent.txt
for wl in C5dis C5nn; do
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:compact -s 6 -c 2 --csv --log-file gpurun_out/r02q_launches_$wl.csv python bench.py --workload $wl --steps 3 --warmup 3 --no-cpu --no-e2e --no-others --strong-rows 0 > /dev/null 2>&1
grep "duration" gpurun_out/r02q_launches_$wl.csv | awk -F'","' '{print $5, $NF}' | cut -c1-60
done
