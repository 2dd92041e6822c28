#!/bin/bash
mkdir -p gpurun_out
for v in 0 1 2; do
  LF_ROUTER_COOP=$v python bench.py --workload c2 --ldd deep --steps 5 --warmup 2 --no-cpu-baseline > gpurun_out/r02_bench8_c2_deep_coop$v.json 2>> gpurun_out/r02_bench8.err
done
LF_ROUTER_COOP=1 python -m pytest tests/test_gpu_kinwave.py -q > gpurun_out/r02_pytest8.log 2>&1
for v in 0 1 2; do echo "coop $v"; tail -1 gpurun_out/r02_bench8_c2_deep_coop$v.json | cut -c1-300; done; tail -2 gpurun_out/r02_pytest8.log; tail -3 gpurun_out/r02_bench8.err
