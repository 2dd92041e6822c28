#!/bin/bash
# 2-GPU box: full single-GPU test suite on GPU 0, the cut-raster checks over NVLink/NCCL, bench lines at N=1 and N=2
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/r02_smi4.txt
nvidia-smi topo -m >> gpurun_out/r02_smi4.txt 2>&1
CUDA_VISIBLE_DEVICES=0 python -m pytest tests -m gpu -q -s --deselect tests/test_gpu_multi.py --deselect tests/test_gpu_bench_configs.py::test_c2_as_named > gpurun_out/r02_pytest4.log 2>&1
NCCL_DEBUG=INFO python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tools/run_dist_check.py > gpurun_out/r02_dist_check_2gpu_nccl.log 2>&1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 tools/run_dist_check.py --stress > gpurun_out/r02_dist_check_2gpu_stress.log 2>&1
python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench4_n1.json 2> gpurun_out/r02_bench4.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02_bench4_n2.json 2>> gpurun_out/r02_bench4.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29614 bench.py --gpus 2 --workload c4 --steps 3 --warmup 1 > gpurun_out/r02_bench4_c4_n2.json 2>> gpurun_out/r02_bench4.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29615 bench.py --gpus 1 --workload c4 --steps 3 --warmup 1 > gpurun_out/r02_bench4_c4_n1.json 2>> gpurun_out/r02_bench4.err
python bench.py --workload c5 --steps 20 --warmup 3 > gpurun_out/r02_bench4_c5.json 2>> gpurun_out/r02_bench4.err
grep -E "passed|failed" gpurun_out/r02_pytest4.log | tail -3; grep -E "^FAILED|^ERROR" gpurun_out/r02_pytest4.log | head; grep -E "DIST CHECK|routing|model " gpurun_out/r02_dist_check_2gpu_nccl.log | cut -c1-300; grep "DIST CHECK" gpurun_out/r02_dist_check_2gpu_stress.log
for f in n1 n2 c4_n2 c4_n1 c5; do echo "== $f"; tail -1 gpurun_out/r02_bench4_$f.json | cut -c1-1800; done; tail -5 gpurun_out/r02_bench4.err
