# One GPU call that re-establishes the evidence of a round (run with: gpurun --timeout 1500 -- 'bash tools/gpu_round_check.sh').
# Outputs land in gpurun_out/; copy what should be judged into profiles/ (named per round).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt 2>&1
# 1. parity: every GPU test, then the driver's smoke
timeout 300 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -25 gpurun_out/pytest_gpu.log | cut -c1-400
timeout 120 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
# 2. the bench line (C3, defaults)
timeout 400 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
cut -c1-1500 gpurun_out/bench.json
# 3. launch list of the same workload (shares of the step; cold cache, serialised)
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_soil|k_of_|k_chan|k_rows|k_u8" -c 1400 --csv \
    --log-file gpurun_out/launches_c3.csv python bench.py --spinup 0 --steps 1 --warmup 1 --no-e2e > gpurun_out/ncu_launch.log 2>&1; echo "ncu launch list rc=$?"
# 4. full capture of the soil stage at C3 size (DRAM traffic per launch -> profiles/rNN_soil_stage_c3_traffic.json):
#    the 11 steps before the captured one launch 3 k_soil kernels each
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_soil" -s 33 -c 3 -f -o gpurun_out/soil_stage_c3_full \
    python bench.py --steps 1 --warmup 0 --no-e2e > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out
