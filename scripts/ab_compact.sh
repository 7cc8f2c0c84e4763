#!/bin/bash
# A/B of the compact (two-pass) path's block width on C5dis.  Usage: gpurun -- bash scripts/ab_compact.sh <tag>
tag=${1:-r02}
mkdir -p gpurun_out
L=$PWD/smmregrid_b200/lib
for v in "" _cw128 _cw64; do
  lib=$L/libsmmregrid_b200$v.so
  for args in "--workload C5dis --steps 10" "--workload C5dis --steps 10 --batch 128" "--workload C5dis --steps 10 --xdtype f64" "--workload C5nn --steps 10 --kernel compact"; do
    SMM_LIB_PATH=$lib bash scripts/ab.sh "lib$v $args" $args
  done
  SMM_LIB_PATH=$lib python -m pytest tests/test_gpu_parity.py -m gpu -q --no-header -k "compact" 2>&1 | tail -1
done | tee gpurun_out/${tag}_ab_compact.txt
