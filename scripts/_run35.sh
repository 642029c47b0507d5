TR4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512"
DFM_BENCH_STEP_TIMES=1 timeout 300 $TR4 bench.py --gpus 4 --steps 20 --warmup 3 --check > gpurun_out/r2_b_n4c.json 2> gpurun_out/r2_b_n4c.err
grep "rank 0. step end times" gpurun_out/r2_b_n4c.err | head -1 | sed 's/.*deltas://' | cut -c1-160
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_b_n4c.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["e2e"]["ms_per_step"], d["value"], d.get("parity"))
PY
