TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
for v in 1 1 1 1; do
DFM_BENCH_NO_SAMPLER=$v DFM_BENCH_STEP_TIMES=1 timeout 300 $TR bench.py --gpus 2 --steps 40 --warmup 3 > gpurun_out/r2_b_n2r.json 2> gpurun_out/r2_b_n2r.err
echo "nosampler=$v $(grep 'rank 0. step end times' gpurun_out/r2_b_n2r.err | head -1 | sed 's/.*deltas://' | cut -c1-210)"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_b_n2r.json").read().strip().splitlines()[-1])
print("   ", d["ms_per_step"], d["e2e"]["ms_per_step"], d["value"])
PY
done
