#!/bin/bash
# pF operator / stress days (the new tests) and the stand-alone soil operators
mkdir -p gpurun_out
python -m pytest tests/test_gpu_soil_ops.py tests/test_capi.py -q -m gpu > gpurun_out/r02_pytest_check19.log 2>&1
tail -12 gpurun_out/r02_pytest_check19.log
