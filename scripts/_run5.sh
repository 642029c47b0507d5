nproc; cat /proc/cpuinfo | grep "model name" | head -1
DFM_BENCH_CPROFILE=gpurun_out/cprof_n2.txt timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_b_n2d.json 2> gpurun_out/r2_b_n2d.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_b_n2d.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["e2e"]["ms_per_step"])
PY
