#!/bin/bash
mkdir -p gpurun_out
for v in 0 1 2 3; do
  LF_CHAN_DEBUG=$v LF_ISO_BPS=6 python bench.py --steps 6 --warmup 2 --spinup 3 --no-e2e > gpurun_out/r02_bench15_dbg$v.json 2>> gpurun_out/r02_bench15.err
done
LF_CHAN_DEBUG=3 LF_ISO_BPS=8 python bench.py --steps 6 --warmup 2 --spinup 3 --no-e2e > gpurun_out/r02_bench15_dbg3_bps8.json 2>> gpurun_out/r02_bench15.err
LF_CHAN_DEBUG=0 LF_ISO_BPS=6 python bench.py --rows 3536 --cols 3536 --steps 6 --warmup 2 --spinup 3 --no-e2e > gpurun_out/r02_bench15_small_dbg0.json 2>> gpurun_out/r02_bench15.err
LF_CHAN_DEBUG=1 LF_ISO_BPS=6 python bench.py --rows 3536 --cols 3536 --steps 6 --warmup 2 --spinup 3 --no-e2e > gpurun_out/r02_bench15_small_dbg1.json 2>> gpurun_out/r02_bench15.err
LF_CHAN_DEBUG=2 LF_ISO_BPS=6 python bench.py --rows 3536 --cols 3536 --steps 6 --warmup 2 --spinup 3 --no-e2e > gpurun_out/r02_bench15_small_dbg2.json 2>> gpurun_out/r02_bench15.err
for f in dbg0 dbg1 dbg2 dbg3 dbg3_bps8 small_dbg0 small_dbg1 small_dbg2; do echo "$f"; tail -1 gpurun_out/r02_bench15_$f.json | cut -c1-260; done; tail -3 gpurun_out/r02_bench15.err
