#!/bin/bash
python scripts/call_latency.py > gpurun_out/r02f_call_latency.txt 2> gpurun_out/r02f_call_latency.err; cat gpurun_out/r02f_call_latency.txt; tail -3 gpurun_out/r02f_call_latency.err
