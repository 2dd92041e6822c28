mkdir -p gpurun_out
timeout 420 python -m pytest tests/test_gpu_model.py -x -q > gpurun_out/pytest_model.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/pytest_model.log
timeout 240 python tools/soil_variants.py --steps 5 > gpurun_out/variants2.log 2>&1; echo "variants rc=$?"
grep variant gpurun_out/variants2.log || tail -20 gpurun_out/variants2.log
