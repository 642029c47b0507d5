TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 200 $TR scripts/check_sharded.py deepfm > gpurun_out/r2_chk2_deepfm.log 2>&1; echo "rc=$?" >> gpurun_out/r2_chk2_deepfm.log
timeout 200 $TR scripts/check_sharded.py xdeepfm_multihot > gpurun_out/r2_chk2_mh.log 2>&1; echo "rc=$?" >> gpurun_out/r2_chk2_mh.log
grep -h "check\|rc=" gpurun_out/r2_chk2_deepfm.log gpurun_out/r2_chk2_mh.log
DFM_BENCH_CPROFILE=gpurun_out/cprof_n2b.txt timeout 300 $TR bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2_b_n2e.json 2> gpurun_out/r2_b_n2e.err
timeout 300 $TR bench.py --gpus 2 --steps 20 --warmup 3 --profile-step gpurun_out/r2_step_n2e.csv > gpurun_out/r2_b_n2f.json 2> gpurun_out/r2_b_n2f.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_b12.json 2> gpurun_out/r2_b12.err
python - <<'PY'
import json
for f in ("r2_b_n2e","r2_b_n2f","r2_b12"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d["ms_per_step"], d["e2e"]["ms_per_step"], d["value"])
    except Exception as e:
        print(f, "failed", e); print(open(f"gpurun_out/{f}.err").read()[-1500:])
PY
