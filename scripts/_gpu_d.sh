#!/bin/bash
# diagnostic: pass 2 of the two-pass path without its stores / without its XT loads / without both
L=$PWD/smmregrid_b200/lib
for v in "" _p2nostore _p2noload _p2neither; do
  for wl in C5dis C5nn; do
    SMM_LIB_PATH=$L/libsmmregrid_b200$v.so ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:compact_apply -s 3 -c 2 --csv --log-file gpurun_out/r02d_p2${v}_$wl.csv python bench.py --workload $wl --steps 3 --warmup 3 --no-cpu --no-e2e --no-others --strong-rows 0 > /dev/null 2>&1
    echo "lib$v $wl: $(grep -c compact_apply gpurun_out/r02d_p2${v}_$wl.csv) rows"
  done
done
