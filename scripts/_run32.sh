TR2="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
TR4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512"
timeout 300 $TR4 bench.py --gpus 4 --steps 20 --warmup 3 > gpurun_out/r2_b_n4b.json 2> gpurun_out/r2_b_n4b.err
timeout 300 $TR2 bench.py --gpus 2 --steps 20 --warmup 3 --check > gpurun_out/r2_b_n2t.json 2> gpurun_out/r2_b_n2t.err
python - <<'PY'
import json
for f in ("r2_b_n4b","r2_b_n2t"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d["ms_per_step"], d["e2e"]["ms_per_step"], d["value"], d.get("parity"))
    except Exception as e:
        print(f, "failed", e); print(open(f"gpurun_out/{f}.err").read()[-2500:])
PY
