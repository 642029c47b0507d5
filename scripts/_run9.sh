TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
for i in 1 2; do
timeout 300 $TR bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2_b_n2j$i.json 2> gpurun_out/r2_b_n2j$i.err
done
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2_b13.json 2> gpurun_out/r2_b13.err
python - <<'PY'
import json
for f in ("r2_b_n2j1","r2_b_n2j2","r2_b13"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d["ms_per_step"], d["e2e"]["ms_per_step"], d["value"], d.get("roofline_bwd",{}).get("ms"))
    except Exception as e:
        print(f, "failed", e); print(open(f"gpurun_out/{f}.err").read()[-1500:])
PY
