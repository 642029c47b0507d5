TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 200 $TR scripts/check_sharded.py deepfm > gpurun_out/r2_chk3_deepfm.log 2>&1; echo "rc=$?" >> gpurun_out/r2_chk3_deepfm.log
timeout 200 $TR scripts/check_sharded.py xdeepfm_multihot > gpurun_out/r2_chk3_mh.log 2>&1; echo "rc=$?" >> gpurun_out/r2_chk3_mh.log
timeout 200 $TR scripts/check_sharded.py deepfm nccl > gpurun_out/r2_chk3_nccl.log 2>&1; echo "rc=$?" >> gpurun_out/r2_chk3_nccl.log
grep -h "check\|rc=\|Error\|error" gpurun_out/r2_chk3_deepfm.log gpurun_out/r2_chk3_mh.log gpurun_out/r2_chk3_nccl.log | cut -c1-250
timeout 300 $TR bench.py --gpus 2 --steps 20 --warmup 3 --profile-step gpurun_out/r2_step_n2g.csv > gpurun_out/r2_b_n2g.json 2> gpurun_out/r2_b_n2g.err
DFM_BENCH_CPROFILE=gpurun_out/cprof_n2c.txt timeout 300 $TR bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2_b_n2h.json 2> gpurun_out/r2_b_n2h.err
timeout 600 python -m pytest tests/test_sharded.py -q -m gpu 2>&1 | tail -3
python - <<'PY'
import json
for f in ("r2_b_n2g","r2_b_n2h"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d["ms_per_step"], d["e2e"]["ms_per_step"], d["value"])
    except Exception as e:
        print(f, "failed", e); print(open(f"gpurun_out/{f}.err").read()[-1500:])
PY
