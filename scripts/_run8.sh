TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
DFM_BENCH_TRACE=gpurun_out/trace_n2b.json DFM_BENCH_TRACE_CUDA_ONLY=1 timeout 300 $TR bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_b_n2i.json 2> gpurun_out/r2_b_n2i.err
gzip -f gpurun_out/trace_n2b.json
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_b_n2i.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["e2e"]["ms_per_step"], d["value"])
PY
