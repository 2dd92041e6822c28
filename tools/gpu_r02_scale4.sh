#!/bin/bash
# the driver's multi-GPU command (N = $1), plus the topology of the box for the host-link figures
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r02_topology_n$N.txt 2>&1
lscpu | grep -E "Model name|Socket|NUMA|^CPU\(s\)" >> gpurun_out/r02_topology_n$N.txt 2>&1
python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node $N --master-port 29662 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02_scale4_c3_n$N.json 2> gpurun_out/r02_scale4_n$N.err
tail -1 gpurun_out/r02_scale4_c3_n$N.json | python -c "
import sys,json
d=json.loads(sys.stdin.read()); c=d['config']
print({k:d.get(k) for k in ('value','ms_per_step','n_gpus','gpu_launches','host_launch_calls','clocks')}, d.get('stage_ms_per_step'))
e=d['e2e']; print('  e2e', e['value'], e.get('ms_per_step'), e.get('host_link'))
print('  c4_cut', d['c4_cut']['value'], d['c4_cut']['ms_per_step'], {k:d['c4_cut']['config'].get(k) for k in ('cut_edges','trunk_pixels')})"
tail -3 gpurun_out/r02_scale4_n$N.err
