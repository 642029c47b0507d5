set -x
python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r2_t9.log
python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "bench rc=$?"
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
python scripts/ncu_target.py && ncu --set full --clock-control none --import-source on -k regex:'embed_fwd|seg2|stitch|dense_stream' -s 10 -c 8 -o gpurun_out/prof_r2 -f python scripts/ncu_target.py > gpurun_out/ncu_full.log 2>&1
python scripts/ncu_gemm3.py && ncu --set full --clock-control none --import-source on -k regex:gemm3_kernel -s 3 -c 3 -o gpurun_out/prof_r2_gemm3 -f python scripts/ncu_gemm3.py > gpurun_out/ncu_gemm3.log 2>&1
tail -3 gpurun_out/r2_t9.log; cat gpurun_out/r2_bench_n1.json | cut -c1-600
