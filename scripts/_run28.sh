TR8="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
DFM_BENCH_STEP_TIMES=1 timeout 300 $TR8 bench.py --gpus 8 --steps 20 --warmup 3 --profile-step gpurun_out/r2_step_n8c.csv > gpurun_out/r2_b_n8e.json 2> gpurun_out/r2_b_n8e.err
DFM_BENCH_STEP_TIMES=1 timeout 300 $TR8 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r2_b_n8f.json 2> gpurun_out/r2_b_n8f.err
timeout 500 $TR8 bench.py --gpus 8 --steps 5 --warmup 3 --workload xdeepfm_criteo_multihot --check > gpurun_out/r2_b_mh_n8b.json 2> gpurun_out/r2_b_mh_n8b.err
grep "rank 0. step end times" gpurun_out/r2_b_n8e.err gpurun_out/r2_b_n8f.err | sed 's/.*deltas://' | cut -c1-130
python - <<'PY'
import json
for f in ("r2_b_n8e","r2_b_n8f","r2_b_mh_n8b"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d["ms_per_step"], d["e2e"]["ms_per_step"], d["value"], d.get("parity"), d.get("clocks"))
    except Exception as e:
        print(f, "failed", e); print(open(f"gpurun_out/{f}.err").read()[-2500:])
PY
