mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_gpu.log
timeout 240 python tools/soil_variants.py --steps 5 --variants 13,13 > gpurun_out/variants3.log 2>&1; echo "variants rc=$?"
grep variant gpurun_out/variants3.log || tail -20 gpurun_out/variants3.log
LF_SOIL_DEF_MB=6 timeout 240 python tools/soil_variants.py --steps 5 --variants 13,13 > gpurun_out/variants3b.log 2>&1; echo "variants rc=$?"
grep variant gpurun_out/variants3b.log || tail -20 gpurun_out/variants3b.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_chan_isolated|k_of_level" -c 2 -f -o gpurun_out/chan_isolated_full python bench.py --rows 4000 --cols 4000 --steps 1 --warmup 0 --no-e2e > gpurun_out/ncu_chan.log 2>&1; echo "ncu rc=$?"
