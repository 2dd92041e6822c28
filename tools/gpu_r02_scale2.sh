#!/bin/bash
# 8-GPU box: the default multi-GPU command again (NUMA-bound ranks, host launch-call counter), then the full model on ONE
# basin (real cut edges in the model at scale) at N = 8 and N = 1
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
$T --nproc-per-node 8 --master-port 29651 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r02_scale2_c3_n8.json 2> gpurun_out/r02_scale2.err
$T --nproc-per-node 8 --master-port 29652 bench.py --gpus 8 --basin single --ldd-noise 0.2 --steps 6 --warmup 2 --no-c4 --no-cpu-baseline > gpurun_out/r02_scale2_c3basin_n8.json 2>> gpurun_out/r02_scale2.err
CUDA_VISIBLE_DEVICES=0 python bench.py --basin single --ldd-noise 0.2 --steps 6 --warmup 2 --no-cpu-baseline > gpurun_out/r02_scale2_c3basin_n1.json 2>> gpurun_out/r02_scale2.err
for f in c3_n8 c3basin_n8 c3basin_n1; do echo "== $f"; tail -1 gpurun_out/r02_scale2_$f.json | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); c=d['config']
    print({k:d.get(k) for k in ('value','ms_per_step','n_gpus','gpu_launches','host_launch_calls')}, 'e2e', d['e2e']['value'], d.get('stage_ms_per_step'), {k:c.get(k) for k in ('cells_per_rank','cut_edges','trunk_pixels','levels_channel','exchange_aborted','host_cores_bound_to_gpu_numa_node')})
    if 'c4_cut' in d: print('  c4_cut', d['c4_cut']['value'], d['c4_cut']['ms_per_step'], {k:d['c4_cut']['config'].get(k) for k in ('cut_edges','trunk_pixels')})
except Exception as e: print('parse error', e)
"; done; tail -5 gpurun_out/r02_scale2.err
