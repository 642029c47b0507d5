TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
for v in 1 1 1 0 0 2; do
DFM_BENCH_MAX_AHEAD=$v DFM_BENCH_STEP_TIMES=1 timeout 300 $TR bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2_b_n2q.json 2> gpurun_out/r2_b_n2q.err
echo "ahead=$v $(grep 'rank 0. step end times' gpurun_out/r2_b_n2q.err | head -1 | sed 's/.*deltas://' | cut -c1-120)"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_b_n2q.json").read().strip().splitlines()[-1])
print("   ", d["ms_per_step"], d["e2e"]["ms_per_step"], d["value"])
PY
done
