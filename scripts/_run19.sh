timeout 600 python -m pytest tests/test_cin_attention_gpu.py tests/test_models_gpu.py tests/test_baseline_shapes_gpu.py -q -m gpu -k "cin or xdeepfm or model" 2>&1 | tail -8 > gpurun_out/r2_cin2.log
timeout 200 python scripts/ncu_cin.py 8192 2>&1 | grep -v Warning >> gpurun_out/r2_cin2.log
timeout 200 python scripts/ncu_cin.py 65536 2>&1 | grep -v Warning >> gpurun_out/r2_cin2.log
timeout 200 python scripts/probe_cin_tc_bwd.py 2>&1 | grep -v Warning | tail -12 >> gpurun_out/r2_cin2.log
cat gpurun_out/r2_cin2.log
