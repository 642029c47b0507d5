TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
for v in 1 2 3; do
DFM_BENCH_STEP_TIMES=1 timeout 300 $TR bench.py --gpus 2 --steps 40 --warmup 3 > gpurun_out/r2_b_n2s.json 2> gpurun_out/r2_b_n2s.err
echo "run $v $(grep 'rank 0. step end times' gpurun_out/r2_b_n2s.err | head -1 | sed 's/.*deltas://' | cut -c1-210)"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_b_n2s.json").read().strip().splitlines()[-1])
print("   ", d["ms_per_step"], d["e2e"]["ms_per_step"], d["value"], d["clocks"])
PY
done
DFM_BENCH_STEP_TIMES=1 timeout 300 python bench.py --steps 40 --warmup 3 --no-cpu-baseline > gpurun_out/r2_b18.json 2> gpurun_out/r2_b18.err
grep 'step end times' gpurun_out/r2_b18.err | head -1 | sed 's/.*deltas://' | cut -c1-210
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_b18.json").read().strip().splitlines()[-1])
print("N=1", d["ms_per_step"], d["e2e"]["ms_per_step"], d["value"], d["clocks"])
PY
