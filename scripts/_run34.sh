TR2="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
for i in 1 2; do
DFM_BENCH_STEP_TIMES=1 timeout 300 $TR2 bench.py --gpus 2 --steps 30 --warmup 3 > gpurun_out/r2_b_n2u$i.json 2> gpurun_out/r2_b_n2u$i.err
grep "rank 0. step end times" gpurun_out/r2_b_n2u$i.err | head -1 | sed 's/.*deltas://' | cut -c1-160
done
python - <<'PY'
import json
for f in ("r2_b_n2u1","r2_b_n2u2"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d["ms_per_step"], d["e2e"]["ms_per_step"], d["value"])
    except Exception as e:
        print(f, "failed", e); print(open(f"gpurun_out/{f}.err").read()[-2500:])
PY
