mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; rc=$?; echo "pytest rc=$rc"
tail -15 gpurun_out/pytest_gpu.log
if [ $rc -eq 0 ]; then
  timeout 400 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
  python -c "
import json;d=json.load(open('gpurun_out/bench.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline_stencil']['frac'], d['stage_ms_per_step'], d['soil_stats'])"
fi
