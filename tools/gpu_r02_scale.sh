#!/bin/bash
# N-GPU box (N = $1, default 8): the driver's scaling commands -- C3 cut over N GPUs with the C4 line attached
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name --format=csv,noheader | sort | uniq -c > gpurun_out/r02_scale_smi_$N.txt
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
$T --nproc-per-node $N --master-port 29641 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02_scale_c3_n$N.json 2> gpurun_out/r02_scale_n$N.err
if [ "$N" = "8" ]; then
  $T --nproc-per-node 4 --master-port 29642 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/r02_scale_c3_n4.json 2>> gpurun_out/r02_scale_n$N.err
  $T --nproc-per-node 8 --master-port 29643 bench.py --gpus 8 --workload c3-replicas --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_scale_c3replicas_n8.json 2>> gpurun_out/r02_scale_n$N.err
  $T --nproc-per-node 8 --master-port 29644 tools/run_dist_check.py > gpurun_out/r02_dist_check_8gpu.log 2>&1
fi
for f in gpurun_out/r02_scale_c3_n$N.json gpurun_out/r02_scale_c3_n4.json gpurun_out/r02_scale_c3replicas_n8.json; do [ -f $f ] && { echo "== $f"; tail -1 $f | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); c=d['config']
    print({k:d.get(k) for k in ('value','ms_per_step','n_gpus','scaling')}, 'e2e', d['e2e']['value'], d.get('stage_ms_per_step'), {k:c.get(k) for k in ('cells_per_rank','cut_edges','exchange_aborted')})
    if 'c4_cut' in d: print('  c4_cut', d['c4_cut']['value'], d['c4_cut']['ms_per_step'], {k:d['c4_cut']['config'].get(k) for k in ('cells_per_rank','cut_edges','trunk_pixels','levels','exchange_aborted')})
except Exception as e: print('parse error', e)
"; }; done; grep -E "DIST CHECK" gpurun_out/r02_dist_check_8gpu.log 2>/dev/null; tail -5 gpurun_out/r02_scale_n$N.err
