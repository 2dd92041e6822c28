#!/bin/bash
# 1 GPU: the cooperative persistent wavefront of the router (deep networks)
mkdir -p gpurun_out
python -m pytest tests/test_gpu_kinwave.py tests/test_gpu_multi.py -q -s > gpurun_out/r02_pytest7.log 2>&1
python bench.py --workload c2 --ldd deep --steps 5 --warmup 2 > gpurun_out/r02_bench7_c2_deep.json 2> gpurun_out/r02_bench7.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29631 bench.py --gpus 1 --workload c4 --steps 3 --warmup 2 > gpurun_out/r02_bench7_c4_n1.json 2>> gpurun_out/r02_bench7.err
grep -E "passed|failed" gpurun_out/r02_pytest7.log | tail -2; grep -E "^FAILED|^ERROR|Error" gpurun_out/r02_pytest7.log | head
for f in c2_deep c4_n1; do echo "== $f"; tail -1 gpurun_out/r02_bench7_$f.json | cut -c1-900; done; tail -5 gpurun_out/r02_bench7.err
