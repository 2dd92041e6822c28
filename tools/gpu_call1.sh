# round-1 GPU call: occupancy variants of k_soil_fused on C3 + one full ncu capture (source counters) of the default
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt 2>&1
timeout 240 python tools/soil_variants.py --variants 0,4,5 --steps 5 > gpurun_out/variants.log 2>&1; echo "variants rc=$?"
grep variant gpurun_out/variants.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_soil_fused" -c 1 -f -o gpurun_out/soil_fused_full python bench.py --rows 4000 --cols 4000 --steps 1 --warmup 0 --no-e2e > gpurun_out/ncu_full.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out
