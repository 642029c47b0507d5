timeout 600 python -m pytest tests/test_cin_attention_gpu.py tests/test_models_gpu.py -q -m gpu -k "cin or xdeepfm or model" 2>&1 | tail -8 > gpurun_out/r2_cin.log
for v in "DFM_CIN_GENERIC=1" "DFM_CIN_X=0"; do
  env $v timeout 200 python scripts/ncu_cin.py 8192 2>&1 | grep -v Warning | sed "s/^/$v /"
  env $v timeout 200 python scripts/ncu_cin.py 65536 2>&1 | grep -v Warning | sed "s/^/$v /"
done >> gpurun_out/r2_cin.log 2>&1
timeout 300 python bench.py --workload xdeepfm_ml --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_xdeepfm_ml_b.json 2> gpurun_out/r2_bench_xdeepfm_ml_b.err
cat gpurun_out/r2_cin.log
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_bench_xdeepfm_ml_b.json").read().strip().splitlines()[-1])
print("xdeepfm_ml", d["ms_per_step"], d["value"])
PY
