#!/bin/bash
# what the driver runs at round end, on one GPU: GPU tests, smoke, the two bench arms
mkdir -p gpurun_out
python -m pytest tests/ -x -q -m gpu > gpurun_out/r02_pytest_final.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke_final.log 2>&1
python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/r02_bench_final_reference.json 2> gpurun_out/r02_bench_final.err
python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/r02_bench_final_n1.json 2>> gpurun_out/r02_bench_final.err
tail -3 gpurun_out/r02_pytest_final.log; tail -2 gpurun_out/r02_smoke_final.log; tail -1 gpurun_out/r02_bench_final_n1.json | cut -c1-400; tail -1 gpurun_out/r02_bench_final_n1.json | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print(d['e2e']['value'], d['roofline']['frac'], d['roofline_fp64_step'], d['stage_ms_per_step'], d['config'].get('host_cores_bound_to_gpu_numa_node'), d['cpu_baseline']['value'])"; tail -3 gpurun_out/r02_bench_final.err
