TR8="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
TR4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512"
timeout 400 $TR8 bench.py --gpus 8 --steps 20 --warmup 3 --check --profile-step gpurun_out/r2_step_n8b.csv > gpurun_out/r2_b_n8c.json 2> gpurun_out/r2_b_n8c.err
timeout 300 $TR8 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r2_b_n8d.json 2> gpurun_out/r2_b_n8d.err
timeout 300 $TR4 bench.py --gpus 4 --steps 20 --warmup 3 --check > gpurun_out/r2_b_n4.json 2> gpurun_out/r2_b_n4.err
python - <<'PY'
import json
for f in ("r2_b_n8c","r2_b_n8d","r2_b_n4"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d["ms_per_step"], d["e2e"]["ms_per_step"], d["value"], d.get("parity"))
    except Exception as e:
        print(f, "failed", e); print(open(f"gpurun_out/{f}.err").read()[-2500:])
PY
