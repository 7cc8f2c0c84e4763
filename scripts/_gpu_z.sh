#!/bin/bash
python -m pytest tests -x -q -m gpu -k "c5 or compact or scattered or random_operator or named_configs" > gpurun_out/r02z_tests.log 2>&1; tail -3 gpurun_out/r02z_tests.log
{
for wl in C5dis C5nn; do
  bash scripts/ab.sh "$wl f32" --workload $wl --steps 20 --warmup 3
  bash scripts/ab.sh "$wl f64 data, compact" --workload $wl --kernel compact --xdtype f64 --steps 10 --warmup 3
done
} > gpurun_out/r02z_ab_compact_pass2.txt 2>&1
cat gpurun_out/r02z_ab_compact_pass2.txt
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:compact -c 4 --csv --log-file gpurun_out/r02z_launches_c5dis.csv python bench.py --workload C5dis --steps 3 --warmup 3 --no-cpu --no-e2e --no-others --strong-rows 0 > /dev/null 2>&1
grep "duration" gpurun_out/r02z_launches_c5dis.csv | cut -d, -f5,15-
