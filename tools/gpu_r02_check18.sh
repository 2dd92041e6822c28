#!/bin/bash
# 2 GPUs, final build: the cut router / model / structures bit-identical to one GPU, then the driver's N = 2 command
mkdir -p gpurun_out
O=gpurun_out
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 $T --nproc-per-node 2 --master-port 29681 tools/run_dist_check.py > $O/r02_dist_check_2gpu_final.log 2>&1
grep -E "PASSED|FAILED" $O/r02_dist_check_2gpu_final.log | tail -2
timeout 600 $T --nproc-per-node 2 --master-port 29683 bench.py --gpus 2 --steps 10 --warmup 3 > $O/r02_check18_c3_n2.json 2> $O/r02_check18.err
tail -1 $O/r02_check18_c3_n2.json | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print({k:d.get(k) for k in ('value','ms_per_step','n_gpus','gpu_launches','clocks')}, d.get('stage_ms_per_step'))
e=d['e2e']; print('  e2e', e['value'], e.get('ms_per_step'), e.get('host_link')); print('  packed', d['e2e_packed']['value'], d['e2e_packed']['ms_per_step'])
print('  c4_cut', d['c4_cut']['value'], d['c4_cut']['ms_per_step'], {k:d['c4_cut']['config'].get(k) for k in ('cut_edges','trunk_pixels')})"
tail -3 $O/r02_check18.err
