set -x
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -3 > gpurun_out/r2_final_tests.log; cat gpurun_out/r2_final_tests.log
( time python bench.py > gpurun_out/r2_bench_n1_final.json 2> gpurun_out/r2_bench_n1_final.err ) 2>&1 | grep real; echo "bench rc=$?"
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
python scripts/ncu_target.py && ncu --set full --clock-control none --import-source on -k regex:'embed_fwd|seg2|stitch|dense_stream' -s 10 -c 8 -o gpurun_out/prof_r2 -f python scripts/ncu_target.py > gpurun_out/ncu_full.log 2>&1
python scripts/ncu_gemm3.py && ncu --set full --clock-control none --import-source on -k regex:gemm3_kernel -s 3 -c 3 -o gpurun_out/prof_r2_gemm3 -f python scripts/ncu_gemm3.py > gpurun_out/ncu_gemm3.log 2>&1
python scripts/ncu_cin.py 8192 > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:cin_tc -s 14 -c 7 -o gpurun_out/prof_r2_cin -f python scripts/ncu_cin.py 8192 > gpurun_out/ncu_cin.log 2>&1
for w in xdeepfm_ml_yaml xdeepfm_ml attention_ml; do
  timeout 300 python bench.py --workload $w --steps 10 --warmup 3 > gpurun_out/r2_bench_$w.json 2> gpurun_out/r2_bench_$w.err; echo "$w rc=$?"
done
python - <<'PY'
import json,glob
for f in ["gpurun_out/r2_bench_n1_final.json"]+sorted(glob.glob("gpurun_out/r2_bench_*ml*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d.get("ms_per_step"), d.get("value"), (d.get("e2e") or {}).get("value"), (d.get("roofline_dnn") or {}).get("frac"), (d.get("roofline_cin") or {}).get("frac"), d.get("clocks"))
    except Exception as e:
        print(f, "failed", e)
PY
