DFM_BENCH_TRACE=gpurun_out/trace_n1.json DFM_BENCH_TRACE_E2E=1 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_b11.json 2> gpurun_out/r2_b11.err
DFM_BENCH_TRACE=gpurun_out/trace_n2.json timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_b_n2c.json 2> gpurun_out/r2_b_n2c.err
gzip -f gpurun_out/trace_n1.json gpurun_out/trace_n2.json
ls -la gpurun_out/*.gz; cut -c1-200 gpurun_out/r2_b_n2c.json; tail -3 gpurun_out/r2_b_n2c.err
