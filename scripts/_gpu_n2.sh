#!/bin/bash
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --impl reference --steps 5 --warmup 3 > gpurun_out/r02f_bench_reference_n2.json 2> gpurun_out/r02f_bench_reference_n2.err; tail -c 300 gpurun_out/r02f_bench_reference_n2.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 > gpurun_out/r02f_bench_c4_n2.json 2> gpurun_out/r02f_bench_c4_n2.err; tail -c 400 gpurun_out/r02f_bench_c4_n2.json; tail -3 gpurun_out/r02f_bench_c4_n2.err
