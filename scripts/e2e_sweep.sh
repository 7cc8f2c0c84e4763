#!/bin/bash
# End-to-end (host buffers in, host buffers out) sweep of the host pipeline's knobs on one GPU:
# chunk size, number of H2D streams, ring depth.  Usage: gpurun -- bash scripts/e2e_sweep.sh <tag>
tag=${1:-r02}
mkdir -p gpurun_out
B="python bench.py --steps 3 --no-cpu --no-others --strong-rows 0 --e2e-steps 10"
run() { # name, env...
  name=$1; shift
  env "$@" $B 2> gpurun_out/${tag}_e2e_${name}.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['e2e']
print('$name', 'pinned %.2f GB/s (%.3f of ceiling %.2f)  pageable %.2f GB/s' % (e['h2d_gbs'], e['frac_of_ceiling'], e['h2d_ceiling_gbs'], e['pageable']['h2d_gbs']))
" | tee -a gpurun_out/${tag}_e2e_sweep.txt
}
: > gpurun_out/${tag}_e2e_sweep.txt
run default SMM_NOP=1
run in1 SMM_HOST_IN_STREAMS=1
run chunk32 SMM_HOST_CHUNK_MB=32
run chunk64 SMM_HOST_CHUNK_MB=64
run chunk256 SMM_HOST_CHUNK_MB=256
run chunk512 SMM_HOST_CHUNK_MB=512
run in1_chunk256 SMM_HOST_IN_STREAMS=1 SMM_HOST_CHUNK_MB=256
run slots2 SMM_HOST_SLOTS=2
run threads16 SMM_HOST_COPY_THREADS=16
run threads4 SMM_HOST_COPY_THREADS=4
