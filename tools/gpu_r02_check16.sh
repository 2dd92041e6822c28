#!/bin/bash
mkdir -p gpurun_out
for v in 5 6 7; do
  LF_ISO_BPS=$v python bench.py --steps 8 --warmup 3 --no-e2e > gpurun_out/r02_bench16_big_$v.json 2>> gpurun_out/r02_bench16.err
  LF_ISO_BPS=$v python bench.py --rows 3536 --cols 3536 --steps 10 --warmup 3 --no-e2e > gpurun_out/r02_bench16_small_$v.json 2>> gpurun_out/r02_bench16.err
done
for v in 5 6 7; do for s in big small; do echo "$s bps $v"; tail -1 gpurun_out/r02_bench16_${s}_$v.json | cut -c1-250; done; done; tail -3 gpurun_out/r02_bench16.err
