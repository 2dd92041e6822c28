mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_gpu.log
timeout 400 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
cut -c1-1800 gpurun_out/bench.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_soil|k_of_|k_chan|k_rows|k_u8" -c 1400 --csv --log-file gpurun_out/launches_c3.csv python bench.py --spinup 0 --steps 1 --warmup 1 --no-e2e > gpurun_out/ncu_launch.log 2>&1; echo "ncu1 rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_soil" -s 33 -c 3 -f -o gpurun_out/soil_stage_c3_full python bench.py --steps 1 --warmup 0 --no-e2e > gpurun_out/ncu_full.log 2>&1; echo "ncu2 rc=$?"
ls -la gpurun_out | head -30
