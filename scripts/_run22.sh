TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -4
timeout 200 $TR scripts/check_sharded.py deepfm > gpurun_out/r2_chk6_deepfm.log 2>&1; echo "rc=$?" >> gpurun_out/r2_chk6_deepfm.log
grep -h "check\|rc=\|Error" gpurun_out/r2_chk6_deepfm.log | cut -c1-300
for i in 1 2; do
timeout 300 $TR bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2_b_n2n$i.json 2> gpurun_out/r2_b_n2n$i.err
done
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python - <<'PY'
import json
for f in ("r2_b_n2n1","r2_b_n2n2"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d["ms_per_step"], d["e2e"]["ms_per_step"], d["value"])
    except Exception as e:
        print(f, "failed", e); print(open(f"gpurun_out/{f}.err").read()[-1500:])
PY
