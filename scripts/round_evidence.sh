#!/bin/bash
# Round evidence on the GPU box (one GPU): tests, both bench arms, the launch list of the bench
# command and `ncu --set full` captures of the dominant kernel of C4 / C2 / C3 / C5dis.
# Usage: gpurun -- bash scripts/round_evidence.sh <tag>     (outputs under gpurun_out/; summaries are
# made here afterwards with scripts/ncu_summary.py and scripts/make_traffic.py and copied to profiles/)
tag=${1:-r02}
mkdir -p gpurun_out
Q="--steps 3 --warmup 3 --no-cpu --no-e2e --no-others --strong-rows 0"
python -m pytest tests -m gpu -q --no-header > gpurun_out/${tag}_pytest.log 2>&1; tail -3 gpurun_out/${tag}_pytest.log
python bench.py --impl reference > gpurun_out/${tag}_bench_reference.json 2> gpurun_out/${tag}_bench_reference.err
python bench.py > gpurun_out/${tag}_bench_c4_n1.json 2> gpurun_out/${tag}_bench_c4_n1.err; tail -c 800 gpurun_out/${tag}_bench_c4_n1.json
python bench.py $Q > gpurun_out/${tag}_plain_c4.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches_c4.csv \
    python bench.py $Q > gpurun_out/${tag}_launches_c4.log 2>&1
# (3 warm-up + 3 timed applies per run: skip the first 3 matching launches, keep the next two --
# for the two-pass compact path that is one launch of each of its kernels)
for wl in C4 C2 C3 C3tri C5dis C5nn; do
  python bench.py --workload $wl $Q > gpurun_out/${tag}_plain_${wl}.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:'staged_kernel|compact|gather_kernel' -s 3 -c $( { [ $wl = C5dis ] || [ $wl = C5nn ]; } && echo 2 || echo 1) -f \
      -o gpurun_out/${tag}_${wl} python bench.py --workload $wl $Q > gpurun_out/${tag}_ncu_${wl}.log 2>&1
done
ls -la gpurun_out | tail -20
