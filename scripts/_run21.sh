for i in 1 2 3; do
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2_b17_$i.json 2> gpurun_out/r2_b17_$i.err
done
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_b17_4.json 2> gpurun_out/r2_b17_4.err
python - <<'PY'
import json
for i in (1,2,3,4):
    d=json.loads(open(f"gpurun_out/r2_b17_{i}.json").read().strip().splitlines()[-1])
    print(i, d["ms_per_step"], d["e2e"]["ms_per_step"], d["value"])
PY
