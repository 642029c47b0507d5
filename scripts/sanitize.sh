#!/bin/bash
# compute-sanitizer passes over the hand-written kernels on small shapes (sanitizers).
# usage (on a GPU box): bash scripts/sanitize.sh <out_dir>;  logs: <out_dir>/sanitizer_{memcheck,racecheck,synccheck}_*.log
out=${1:-gpurun_out}
CS="compute-sanitizer --print-limit 20 --error-exitcode 99"
run() {   # tool tag pytest-args...
  tool=$1; tag=$2; shift 2
  timeout 900 $CS --tool $tool python -m pytest -q -x -m gpu "$@" > $out/sanitizer_${tool}_${tag}.log 2>&1
  echo "$tool $tag rc=$?" | tee -a $out/sanitizer_summary.log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" $out/sanitizer_${tool}_${tag}.log | tail -3 | tee -a $out/sanitizer_summary.log
}
: > $out/sanitizer_summary.log
# K1 / K2 (gather + pool + FM, sort + segmented reduce + stitch), every field kind and combiner, long segments
run memcheck embed tests/test_embedding_gpu.py -k "three_views or reference_autograd or long_segments or keys_sort or empty_batch or row_sparse"
run racecheck embed tests/test_embedding_gpu.py -k "three_views or reference_autograd or long_segments"
# tcgen05 CIN (forward, backward-data, dW), fp32 CIN, attention
run memcheck cin_attn tests/test_cin_attention_gpu.py -k "cin_golden or tf32_vs_oracle or own_relu or attention_golden"
run racecheck cin_attn tests/test_cin_attention_gpu.py -k "cin_golden or own_relu or attention_golden"
# DNN tower: tcgen05 3xTF32 GEMM (TMA loads, TMA-store epilogue), fused BN / activation passes
run memcheck dnn tests/test_dnn_gpu.py -k "gemm3 or matches_eager"
run racecheck dnn tests/test_dnn_gpu.py -k "gemm3"
# sharded path on emulated ranks (route / gather / pack / owner backward) + row-sparse Adam
run memcheck shard tests/test_sharded.py tests/test_optim_gpu.py
run synccheck embed tests/test_embedding_gpu.py -k "reference_autograd or long_segments"
cat $out/sanitizer_summary.log
