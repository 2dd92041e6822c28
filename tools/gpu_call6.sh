mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_gpu.log
timeout 240 python tools/soil_variants.py --steps 5 --variants 13,13 > gpurun_out/variants4.log 2>&1; echo "variants rc=$?"
grep variant gpurun_out/variants4.log || tail -20 gpurun_out/variants4.log
LF_SOIL_DEF_MB=8 timeout 240 python tools/soil_variants.py --steps 5 --variants 13,13 > gpurun_out/variants4b.log 2>&1; echo "variants rc=$?"
grep variant gpurun_out/variants4b.log || tail -20 gpurun_out/variants4b.log
