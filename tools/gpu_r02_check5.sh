#!/bin/bash
# 2-GPU box: structures on a cut raster, graph replay in the router, single-basin workloads with real cut edges
mkdir -p gpurun_out
CUDA_VISIBLE_DEVICES=0 python -m pytest tests -m gpu -q --deselect tests/test_gpu_multi.py --deselect tests/test_gpu_bench_configs.py::test_c2_as_named > gpurun_out/r02_pytest5.log 2>&1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29621 tools/run_dist_check.py > gpurun_out/r02_dist_check5.log 2>&1
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
$T --nproc-per-node 2 --master-port 29622 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02_bench5_n2.json 2> gpurun_out/r02_bench5.err
$T --nproc-per-node 1 --master-port 29623 bench.py --gpus 1 --workload c4 --steps 3 --warmup 2 > gpurun_out/r02_bench5_c4_n1.json 2>> gpurun_out/r02_bench5.err
python bench.py --basin single --steps 5 --warmup 2 --no-cpu-baseline > gpurun_out/r02_bench5_c3single_n1.json 2>> gpurun_out/r02_bench5.err
$T --nproc-per-node 2 --master-port 29624 bench.py --gpus 2 --basin single --steps 5 --warmup 2 --no-c4 --no-cpu-baseline > gpurun_out/r02_bench5_c3single_n2.json 2>> gpurun_out/r02_bench5.err
python bench.py --workload c5 --steps 30 --warmup 3 > gpurun_out/r02_bench5_c5_n1.json 2>> gpurun_out/r02_bench5.err
$T --nproc-per-node 2 --master-port 29625 bench.py --gpus 2 --workload c5 --steps 30 --warmup 3 > gpurun_out/r02_bench5_c5_n2.json 2>> gpurun_out/r02_bench5.err
grep -E "passed|failed" gpurun_out/r02_pytest5.log | tail -2; grep -E "^FAILED|^ERROR" gpurun_out/r02_pytest5.log | head; grep -E "DIST CHECK|model " gpurun_out/r02_dist_check5.log | cut -c1-260
for f in n2 c4_n1 c3single_n1 c3single_n2 c5_n1 c5_n2; do echo "== $f"; tail -1 gpurun_out/r02_bench5_$f.json | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read())
    c=d['config']; print({k:d.get(k) for k in ('value','ms_per_step','n_gpus','gpu_launches','kernels_per_step','without_cuda_graphs')}, 'e2e', d['e2e']['value'], {k:c.get(k) for k in ('cells_per_rank','cut_edges','trunk_pixels','levels','levels_channel','exchange_aborted')}, d.get('stage_ms_per_step'))
    if 'c4_cut' in d: print('  c4_cut', d['c4_cut']['value'], d['c4_cut']['ms_per_step'], {k:d['c4_cut']['config'].get(k) for k in ('cells_per_rank','cut_edges','trunk_pixels','levels','exchange_aborted')})
except Exception as e: print('parse error', e)
"; done; tail -8 gpurun_out/r02_bench5.err
