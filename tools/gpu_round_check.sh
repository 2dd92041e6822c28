mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_gpu.log
timeout 400 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
cat gpurun_out/bench.json | cut -c1-3000
timeout 300 python tools/soil_variants.py --steps 5 > gpurun_out/variants.log 2>&1; echo "variants rc=$?"
grep variant gpurun_out/variants.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_soil|k_of_|k_chan" -c 1300 --csv --log-file gpurun_out/launches_c3.csv python bench.py --steps 1 --warmup 1 --no-e2e > gpurun_out/ncu_launch.log 2>&1; echo "ncu1 rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_soil_fused|k_soil_pixel_flagged|k_soil_veg_deferred" -c 3 -o gpurun_out/soil_fused_full python bench.py --rows 4000 --cols 4000 --steps 1 --warmup 0 --no-e2e > gpurun_out/ncu_full.log 2>&1; echo "ncu2 rc=$?"
ls -la gpurun_out
