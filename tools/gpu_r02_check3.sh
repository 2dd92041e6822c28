#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -s --deselect tests/test_gpu_multi.py 2>&1 | tail -150 > gpurun_out/r02_pytest3.log
python -m pytest tests/test_gpu_multi.py -q -s 2>&1 | tail -80 > gpurun_out/r02_pytest_multi_1gpu.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench3.json 2> gpurun_out/r02_bench3.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench3_reference.json 2>> gpurun_out/r02_bench3.err
ncu --metrics sm__inst_executed_pipe_fp64.sum,sm__inst_executed.sum,gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"k_chan_isolated_ws|k_feeder|k_soil_staged|k_soil_veg_deferred|k_soil_pixel_flagged|k_of_level" -c 60 --csv --log-file gpurun_out/r02_ncu_metrics_4000.csv python bench.py --rows 4000 --cols 4000 --steps 1 --warmup 1 --spinup 2 --no-e2e > gpurun_out/r02_ncu_metrics.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/r02_launches_c3_10000.csv python bench.py --steps 2 --warmup 1 --spinup 2 --no-e2e > gpurun_out/r02_ncu_launches.log 2>&1
grep -E "passed|failed|FAILED|ERROR" gpurun_out/r02_pytest3.log | tail -30; tail -15 gpurun_out/r02_pytest_multi_1gpu.log | cut -c1-400; cat gpurun_out/r02_bench3.json | cut -c1-3000; tail -3 gpurun_out/r02_bench3.err
