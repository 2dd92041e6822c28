#!/bin/bash
mkdir -p gpurun_out
python tools/debug_adv.py > gpurun_out/r02_debug_adv.log 2>&1
python -m pytest tests -m gpu -x -q -s --deselect tests/test_gpu_bench_configs.py::test_c2_as_named --deselect tests/test_gpu_multi.py 2>&1 | tail -40 > gpurun_out/r02_pytest2.log
python -m pytest tests/test_gpu_multi.py -x -q -s 2>&1 | tail -60 > gpurun_out/r02_pytest_multi_1gpu.log
python bench.py --steps 6 --warmup 2 --no-e2e > gpurun_out/r02_bench2_serial.json 2>> gpurun_out/r02_bench2.err
for b in 2 4 8; do
  LF_EARLY_BPS=$b python bench.py --steps 6 --warmup 2 --no-e2e --overlap 1 > gpurun_out/r02_bench2_bps$b.json 2>> gpurun_out/r02_bench2.err
done
head -50 gpurun_out/r02_debug_adv.log; cat gpurun_out/r02_bench2_*.json; tail -30 gpurun_out/r02_pytest_multi_1gpu.log; tail -12 gpurun_out/r02_pytest2.log; tail -5 gpurun_out/r02_bench2.err
