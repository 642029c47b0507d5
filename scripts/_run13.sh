TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
DFM_BENCH_CPROFILE=gpurun_out/cprof_n2d.txt timeout 300 $TR bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2_b_n2m.json 2> gpurun_out/r2_b_n2m.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_b_n2m.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["e2e"]["ms_per_step"], d["value"])
PY
