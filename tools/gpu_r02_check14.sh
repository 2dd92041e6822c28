#!/bin/bash
# packed int16 forcing / float32 output: the new tests, then the default bench line with the e2e_packed leg
mkdir -p gpurun_out
python -m pytest tests/test_gpu_feeders.py tests/test_gpu_output_io.py -q -m gpu > gpurun_out/r02_pytest_check14.log 2>&1
tail -15 gpurun_out/r02_pytest_check14.log
python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_check14_bench_n1.json 2> gpurun_out/r02_check14_bench.err
tail -1 gpurun_out/r02_check14_bench_n1.json | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['clocks']); print(d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['host_link']); print(d['e2e_packed'])"
tail -3 gpurun_out/r02_check14_bench.err
