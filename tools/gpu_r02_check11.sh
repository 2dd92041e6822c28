#!/bin/bash
mkdir -p gpurun_out
for v in 8 6 0; do
  LF_ISO_BPS=$v python bench.py --steps 6 --warmup 2 --no-e2e > gpurun_out/r02_bench12_iso$v.json 2>> gpurun_out/r02_bench12.err
done
for v in 8 6 0; do echo "iso bps $v"; tail -1 gpurun_out/r02_bench12_iso$v.json | cut -c1-330; done; tail -3 gpurun_out/r02_bench12.err
