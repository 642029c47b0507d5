set -x
python bench.py > gpurun_out/r2_bench_n1_final.json 2> gpurun_out/r2_bench_n1_final.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err; echo "ref rc=$?"
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
python scripts/ncu_target.py && ncu --set full --clock-control none --import-source on -k regex:'embed_fwd|seg2|stitch|dense_stream' -s 10 -c 8 -o gpurun_out/prof_r2 -f python scripts/ncu_target.py > gpurun_out/ncu_full.log 2>&1
python scripts/ncu_gemm3.py && ncu --set full --clock-control none --import-source on -k regex:gemm3_kernel -s 3 -c 3 -o gpurun_out/prof_r2_gemm3 -f python scripts/ncu_gemm3.py > gpurun_out/ncu_gemm3.log 2>&1
for w in deepfm_ml_b4096 xdeepfm_ml_yaml_b4096 xdeepfm_ml_yaml xdeepfm_ml attention_ml; do
  timeout 300 python bench.py --workload $w --steps 10 --warmup 3 > gpurun_out/r2_bench_$w.json 2> gpurun_out/r2_bench_$w.err; echo "$w rc=$?"
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2_bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d.get("ms_per_step"), d.get("value"), (d.get("e2e") or {}).get("value"), (d.get("cpu_baseline") or {}).get("value"), (d.get("cpu_baseline") or {}).get("kind"))
    except Exception as e:
        print(f, "failed", e)
PY
