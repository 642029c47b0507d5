TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 200 $TR scripts/check_sharded.py deepfm > gpurun_out/r2_chk5_deepfm.log 2>&1; echo "rc=$?" >> gpurun_out/r2_chk5_deepfm.log
timeout 200 $TR scripts/check_sharded.py xdeepfm_multihot > gpurun_out/r2_chk5_mh.log 2>&1; echo "rc=$?" >> gpurun_out/r2_chk5_mh.log
grep -h "check\|rc=\|Error\|error" gpurun_out/r2_chk5_deepfm.log gpurun_out/r2_chk5_mh.log | cut -c1-400
for v in 1 0; do
DFM_SHARD_P2P_IDS=$v timeout 300 $TR bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2_b_n2l$v.json 2> gpurun_out/r2_b_n2l$v.err
done
python - <<'PY'
import json
for f in ("r2_b_n2l1","r2_b_n2l0"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d["ms_per_step"], d["e2e"]["ms_per_step"], d["value"])
    except Exception as e:
        print(f, "failed", e); print(open(f"gpurun_out/{f}.err").read()[-2500:])
PY
