#!/bin/bash
# Round evidence on the GPU box: tests, hypothesis soak, both bench arms, ncu launch list.
# Usage: gpurun -- bash scripts/round_evidence.sh   (outputs under gpurun_out/, copied to profiles/ by hand)
mkdir -p gpurun_out
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
SMM_HYP_EXAMPLES=150 python -m pytest tests/test_gpu_properties.py -m gpu -x -q 2>&1 | tail -3
python bench.py --impl reference > gpurun_out/r01i_bench_reference.json 2> gpurun_out/r01i_bench_reference.err; tail -c 600 gpurun_out/r01i_bench_reference.json
python bench.py > gpurun_out/r01i_bench_c4_n1.json 2> gpurun_out/r01i_bench_c4_n1.err; tail -c 1500 gpurun_out/r01i_bench_c4_n1.json
python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/c4_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01i_launches_c4.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/c4_launches.log 2>&1
