#!/bin/bash
# N GPUs, final build, C3 only (no attached C4 line): device-resident, e2e, host-link floor and e2e_packed
N=${1:-8}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node $N --master-port 29691 bench.py --gpus $N --steps 10 --warmup 3 --no-c4 --no-cpu-baseline > gpurun_out/r02_scale5_c3_n$N.json 2> gpurun_out/r02_scale5_n$N.err
tail -1 gpurun_out/r02_scale5_c3_n$N.json | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print({k:d.get(k) for k in ('value','ms_per_step','n_gpus','gpu_launches','clocks')}, d.get('stage_ms_per_step'))
e=d['e2e']; print('  e2e', e['value'], e.get('ms_per_step'), e.get('host_link')); print('  packed', d['e2e_packed']['value'], d['e2e_packed']['ms_per_step'])"
tail -3 gpurun_out/r02_scale5_n$N.err
