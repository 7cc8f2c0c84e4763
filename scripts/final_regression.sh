#!/bin/bash
# final single-GPU regression of the round: smoke, full GPU suite, both bench arms
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02f_smoke.log 2>&1; tail -2 gpurun_out/r02f_smoke.log
python -m pytest tests -m gpu -q --no-header > gpurun_out/r02f_pytest.log 2>&1; tail -3 gpurun_out/r02f_pytest.log
python bench.py --impl reference > gpurun_out/r02f_bench_reference.json 2> gpurun_out/r02f_bench_reference.err
python bench.py > gpurun_out/r02f_bench_c4_n1.json 2> gpurun_out/r02f_bench_c4_n1.err; tail -c 600 gpurun_out/r02f_bench_c4_n1.json
