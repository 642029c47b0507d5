timeout 300 python -m pytest tests/test_dnn_gpu.py -q -m gpu 2>&1 | tail -5 > gpurun_out/r2_g3_tma.log
for v in "DFM_G3_TMA_STORE=0" "DFM_G3_TMA_STORE=1"; do
  env $v timeout 120 python scripts/perf_gemm3.py 2>&1 | grep -v Warning
done >> gpurun_out/r2_g3_tma.log 2>&1
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_b10.json 2> gpurun_out/r2_b10.err
cat gpurun_out/r2_g3_tma.log; python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_b10.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["e2e"], d.get("roofline_bwd",{}).get("ms"), d.get("roofline_path",{}).get("frac"))
PY
