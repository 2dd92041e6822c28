#!/bin/bash
# 1-GPU box: all GPU tests (incl. C2 as named), bench default, soil-stage DRAM traffic at C3 size, C5
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -s > gpurun_out/r02_pytest6.log 2>&1
python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench6_n1.json 2> gpurun_out/r02_bench6.err
LF_SOIL_SIDE=0 python bench.py --steps 6 --warmup 2 --no-e2e > gpurun_out/r02_bench6_noside.json 2>> gpurun_out/r02_bench6.err
python bench.py --workload c5 --steps 30 --warmup 3 > gpurun_out/r02_bench6_c5_n1.json 2>> gpurun_out/r02_bench6.err
python bench.py --workload c2 --ldd deep --steps 5 --warmup 2 > gpurun_out/r02_bench6_c2_deep.json 2>> gpurun_out/r02_bench6.err
python bench.py --workload c2 --ldd shallow --steps 5 --warmup 2 > gpurun_out/r02_bench6_c2_shallow.json 2>> gpurun_out/r02_bench6.err
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__inst_executed_pipe_fp64.sum,sm__inst_executed.sum --clock-control none -k regex:"k_soil|k_feeder|k_chan_isolated_ws|k_chan_post|k_of_post" -c 40 --csv --log-file gpurun_out/r02_ncu_c3_traffic.csv python bench.py --steps 1 --warmup 1 --spinup 10 --no-e2e > gpurun_out/r02_ncu_c3_traffic.log 2>&1
grep -E "passed|failed" gpurun_out/r02_pytest6.log | tail -2; grep -E "^FAILED|^ERROR|C3 bench-data|C2 (deep|shallow)|worst rel|warm-start" gpurun_out/r02_pytest6.log | cut -c1-600 | head -20
for f in n1 noside c5_n1 c2_deep c2_shallow; do echo "== $f"; tail -1 gpurun_out/r02_bench6_$f.json | cut -c1-1200; done; tail -5 gpurun_out/r02_bench6.err
