#!/bin/bash
# TMA pass 1 of the compact path: tests, then A/B
python -m pytest tests -x -q -m gpu -k "c5 or compact or scattered or real_3d or weights_from_files or real_netcdf4 or named_configs" > gpurun_out/r02v_tests.log 2>&1; tail -3 gpurun_out/r02v_tests.log
{
for wl in C5dis C5nn; do
  for k in compact gather auto; do
    for bulk in 1 0; do
      if [ $k != compact ] && [ $bulk = 0 ]; then continue; fi
      SMM_COMPACT_BULK=$bulk bash scripts/ab.sh "$wl kernel=$k bulk=$bulk" --workload $wl --kernel $k --steps 20 --warmup 3
    done
  done
done
for wl in C5dis C5nn; do
  SMM_COMPACT_BULK=1 bash scripts/ab.sh "$wl f64 in kernel=compact bulk=1" --workload $wl --kernel compact --xdtype f64 --steps 10 --warmup 3
  SMM_COMPACT_BULK=0 bash scripts/ab.sh "$wl f64 in kernel=compact bulk=0" --workload $wl --kernel compact --xdtype f64 --steps 10 --warmup 3
  bash scripts/ab.sh "$wl f64 in kernel=gather" --workload $wl --kernel gather --xdtype f64 --steps 10 --warmup 3
done
} > gpurun_out/r02v_ab_compact_tma.txt 2>&1
cat gpurun_out/r02v_ab_compact_tma.txt
