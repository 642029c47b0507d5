TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
for i in 1 2 3 4; do
DFM_BENCH_STEP_TIMES=1 timeout 300 $TR bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2_b_n2p$i.json 2> gpurun_out/r2_b_n2p$i.err
grep "rank 0. step end times" gpurun_out/r2_b_n2p$i.err | head -1 | sed 's/.*deltas://' | cut -c1-200
done
python - <<'PY'
import json
for i in (1,2,3,4):
    d=json.loads(open(f"gpurun_out/r2_b_n2p{i}.json").read().strip().splitlines()[-1])
    print(i, d["ms_per_step"], d["e2e"]["ms_per_step"], d["value"])
PY
