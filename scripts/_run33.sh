timeout 600 python -m pytest tests/test_dnn_gpu.py tests/test_models_gpu.py tests/test_baseline_shapes_gpu.py -q -m gpu 2>&1 | tail -6
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2_b19.json 2> gpurun_out/r2_b19.err
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/r2_b19.json").read().strip().splitlines()[-1])
    print(d["ms_per_step"], d["e2e"]["ms_per_step"], d["value"], d["gpu_launches"])
except Exception as e:
    print("failed", e); print(open("gpurun_out/r2_b19.err").read()[-2000:])
PY
