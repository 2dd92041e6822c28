mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -25 gpurun_out/pytest_gpu.log | cut -c1-400
timeout 120 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_soil|k_of_|k_chan|k_rows|k_u8" -c 1400 --csv --log-file gpurun_out/launches_c3_final.csv python bench.py --spinup 0 --steps 1 --warmup 1 --no-e2e > gpurun_out/ncu_launch.log 2>&1; echo "ncu1 rc=$?"
