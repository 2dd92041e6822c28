#!/bin/bash
# narrow-run kernels: the whole GPU suite, smoke, then the workloads they matter for and the driver's default lines
mkdir -p gpurun_out
O=gpurun_out
python -m pytest tests/ -q -m gpu > $O/r02_pytest_check17.log 2>&1
tail -5 $O/r02_pytest_check17.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/r02_smoke_check17.log 2>&1; tail -1 $O/r02_smoke_check17.log
python bench.py --workload c2 --ldd deep --steps 5 --warmup 2 --no-cpu-baseline > $O/r02_check17_c2_deep.json 2> $O/r02_check17.err
LF_ROUTER_NARROW=0 python bench.py --workload c2 --ldd deep --steps 5 --warmup 2 --no-cpu-baseline > $O/r02_check17_c2_deep_narrow0.json 2>> $O/r02_check17.err
python bench.py --workload c2 --ldd shallow --steps 5 --warmup 2 --no-cpu-baseline > $O/r02_check17_c2_shallow.json 2>> $O/r02_check17.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29671 bench.py --gpus 1 --workload c4 --steps 3 --warmup 2 > $O/r02_check17_c4_n1.json 2>> $O/r02_check17.err
python bench.py --basin single --ldd-noise 0.2 --steps 6 --warmup 2 --no-cpu-baseline > $O/r02_check17_c3basin_n1.json 2>> $O/r02_check17.err
python bench.py --workload c5 --steps 30 --warmup 3 > $O/r02_check17_c5_n1.json 2>> $O/r02_check17.err
python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 > $O/r02_check17_reference.json 2>> $O/r02_check17.err
python bench.py --gpus 1 --steps 10 --warmup 3 > $O/r02_check17_c3_n1.json 2>> $O/r02_check17.err
for f in c2_deep c2_deep_narrow0 c2_shallow c4_n1 c3basin_n1 c5_n1 c3_n1; do
  tail -1 $O/r02_check17_$f.json | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('$f', d.get('value'), d.get('ms_per_step'), d.get('gpu_launches'), (d.get('e2e') or {}).get('value'), d.get('stage_ms_per_step'))"
done
tail -1 $O/r02_check17_reference.json | cut -c1-300
tail -3 $O/r02_check17.err
