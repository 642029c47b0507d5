TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
timeout 400 $TR bench.py --gpus 8 --steps 20 --warmup 3 --check --profile-step gpurun_out/r2_step_n8.csv > gpurun_out/r2_b_n8.json 2> gpurun_out/r2_b_n8.err
DFM_BENCH_TRACE=gpurun_out/trace_n8.json DFM_BENCH_TRACE_CUDA_ONLY=1 timeout 300 $TR bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r2_b_n8b.json 2> gpurun_out/r2_b_n8b.err
gzip -f gpurun_out/trace_n8.json
timeout 600 $TR bench.py --gpus 8 --steps 5 --warmup 3 --workload xdeepfm_criteo_multihot --check > gpurun_out/r2_b_mh_n8.json 2> gpurun_out/r2_b_mh_n8.err
python - <<'PY'
import json
for f in ("r2_b_n8","r2_b_n8b","r2_b_mh_n8"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d["ms_per_step"], d["e2e"]["ms_per_step"], d["value"], d.get("parity"))
    except Exception as e:
        print(f, "failed", e); print(open(f"gpurun_out/{f}.err").read()[-2500:])
PY
