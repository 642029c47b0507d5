TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 200 $TR scripts/check_sharded.py deepfm > gpurun_out/r2_chk4_deepfm.log 2>&1; echo "rc=$?" >> gpurun_out/r2_chk4_deepfm.log
grep -h "check\|rc=\|Error" gpurun_out/r2_chk4_deepfm.log | cut -c1-200
for v in "-1" "0" "1"; do
DFM_BENCH_MAX_AHEAD=$v timeout 300 $TR bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2_b_n2k$v.json 2> gpurun_out/r2_b_n2k$v.err
done
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2_b14.json 2> gpurun_out/r2_b14.err
timeout 600 python -m pytest tests -q -m gpu -x 2>&1 | tail -3
python - <<'PY'
import json
for f in ("r2_b_n2k-1","r2_b_n2k0","r2_b_n2k1","r2_b14"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d["ms_per_step"], d["e2e"]["ms_per_step"], d["value"])
    except Exception as e:
        print(f, "failed", e); print(open(f"gpurun_out/{f}.err").read()[-1500:])
PY
