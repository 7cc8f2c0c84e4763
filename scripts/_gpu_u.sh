#!/bin/bash
python -m pytest tests/test_gpu_round2.py -x -q -m gpu -k "weights_from_files or real_3d or real_netcdf4" > gpurun_out/r02u_tests.log 2>&1; tail -5 gpurun_out/r02u_tests.log
