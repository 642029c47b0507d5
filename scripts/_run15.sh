timeout 300 python -m pytest tests/test_dnn_gpu.py -q -m gpu 2>&1 | tail -5 > gpurun_out/r2_g3_swap.log
for v in "DFM_G3_DW_SWAP=0" "DFM_G3_DW_SWAP=1"; do
  env $v timeout 120 python scripts/perf_gemm3.py 2>&1 | grep -v Warning
done >> gpurun_out/r2_g3_swap.log 2>&1
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2_b15.json 2> gpurun_out/r2_b15.err
cat gpurun_out/r2_g3_swap.log; python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_b15.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["e2e"]["ms_per_step"], d["value"], d.get("roofline_bwd",{}).get("ms"))
PY
