#!/bin/bash
python -m pytest tests -x -q -m gpu -k "c5 or compact or scattered or real_3d or named_configs" > gpurun_out/r02w_tests.log 2>&1; tail -3 gpurun_out/r02w_tests.log
{
for wl in C5dis C5nn; do
  bash scripts/ab.sh "$wl kernel=compact" --workload $wl --kernel compact --steps 20 --warmup 3
  SMM_COMPACT_BULK=0 bash scripts/ab.sh "$wl kernel=compact bulk=0" --workload $wl --kernel compact --steps 20 --warmup 3
  bash scripts/ab.sh "$wl f64 in kernel=compact rows=32" --workload $wl --kernel compact --xdtype f64 --steps 10 --warmup 3
  SMM_COMPACT_F64_ROWS=0 bash scripts/ab.sh "$wl f64 in kernel=compact rows=64" --workload $wl --kernel compact --xdtype f64 --steps 10 --warmup 3
  SMM_COMPACT_BULK=0 bash scripts/ab.sh "$wl f64 in kernel=compact bulk=0" --workload $wl --kernel compact --xdtype f64 --steps 10 --warmup 3
done
} > gpurun_out/r02w_ab_compact_tma.txt 2>&1
cat gpurun_out/r02w_ab_compact_tma.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 30 --csv --log-file gpurun_out/r02w_launches_c5dis.csv python bench.py --workload C5dis --kernel compact --steps 3 --warmup 3 --no-cpu --no-e2e --no-others --strong-rows 0 > /dev/null 2>&1
grep -c . gpurun_out/r02w_launches_c5dis.csv
