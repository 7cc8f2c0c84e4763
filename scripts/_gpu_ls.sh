#!/bin/bash
python -m pytest tests/test_gpu_round2.py -x -q -m gpu -k "level_partition or compact_first_pass" > gpurun_out/r02g_tests.log 2>&1; tail -3 gpurun_out/r02g_tests.log
python scripts/level_shard_2gpu.py > gpurun_out/r02g_level_shard_n1.txt 2>&1; tail -6 gpurun_out/r02g_level_shard_n1.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 scripts/level_shard_2gpu.py > gpurun_out/r02g_level_shard_n2.txt 2>&1; tail -8 gpurun_out/r02g_level_shard_n2.txt
