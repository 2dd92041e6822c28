mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_chan_isolated" -c 1 -f -o gpurun_out/chan_isolated_full python bench.py --rows 4000 --cols 4000 --steps 1 --warmup 0 --no-e2e > gpurun_out/ncu_chan.log 2>&1; echo "ncu rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_chan_diagonal" -s 200 -c 1 -f -o gpurun_out/chan_diagonal_full python bench.py --rows 4000 --cols 4000 --steps 1 --warmup 0 --no-e2e > gpurun_out/ncu_chan2.log 2>&1; echo "ncu rc=$?"
