timeout 300 python -m pytest tests/test_dnn_gpu.py tests/test_models_gpu.py -q -m gpu 2>&1 | tail -3
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --profile-step gpurun_out/r2_step_n1b.csv > gpurun_out/r2_b16.json 2> gpurun_out/r2_b16.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_b16.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["e2e"]["ms_per_step"], d["value"])
PY
grep "tw::" gpurun_out/r2_step_n1b.csv | cut -c1-110
