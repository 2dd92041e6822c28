mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_gpu.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_soil_staged" -c 1 -f -o gpurun_out/soil_staged_full python bench.py --rows 4000 --cols 4000 --steps 1 --warmup 0 --no-e2e > gpurun_out/ncu_staged.log 2>&1; echo "ncu rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_soil_veg_deferred|k_soil_pixel_flagged" -c 7 -f -o gpurun_out/soil_deferred_full python bench.py --rows 4000 --cols 4000 --steps 1 --warmup 0 --no-e2e > gpurun_out/ncu_deferred.log 2>&1; echo "ncu2 rc=$?"
timeout 400 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
cut -c1-2500 gpurun_out/bench.json
