#!/bin/bash
SMM_COMPACT_ORDER=dst python -m pytest tests -x -q -m gpu -k "c5 or compact or scattered" > gpurun_out/r02y_tests_dst.log 2>&1; tail -3 gpurun_out/r02y_tests_dst.log
{
for wl in C5dis C5nn; do
  for ord in src dst; do
    SMM_COMPACT_ORDER=$ord bash scripts/ab.sh "$wl f32 XT order=$ord" --workload $wl --steps 20 --warmup 3
    SMM_COMPACT_ORDER=$ord bash scripts/ab.sh "$wl f64 data, compact, XT order=$ord" --workload $wl --kernel compact --xdtype f64 --steps 10 --warmup 3
  done
done
} > gpurun_out/r02y_ab_compact_xt_order.txt 2>&1
cat gpurun_out/r02y_ab_compact_xt_order.txt
SMM_COMPACT_ORDER=dst ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:compact -c 8 --csv --log-file gpurun_out/r02y_launches_c5dis_dst.csv python bench.py --workload C5dis --steps 3 --warmup 3 --no-cpu --no-e2e --no-others --strong-rows 0 > /dev/null 2>&1
grep -c . gpurun_out/r02y_launches_c5dis_dst.csv
