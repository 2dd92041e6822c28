#!/bin/bash
# round 2, first GPU check: all GPU tests, the C3 bench line, early-launch footprint sweep
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -s 2>&1 | tail -40 > gpurun_out/r02_pytest1.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench1.json 2> gpurun_out/r02_bench1.err
for b in 1 3 4 6; do
  LF_EARLY_BPS=$b python bench.py --steps 6 --warmup 2 --no-e2e > gpurun_out/r02_bench1_bps$b.json 2>> gpurun_out/r02_bench1.err
done
tail -5 gpurun_out/r02_pytest1.log; cat gpurun_out/r02_bench1.json | head -c 3000; cat gpurun_out/r02_bench1_bps*.json
