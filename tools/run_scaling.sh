#!/bin/bash
# Multi-GPU evidence for one node size N (run under `gpurun --gpus N`): bench.py (weak scaling, no collective), layer-sharded
# 70B-shape fp4 weight quantization, tensor-parallel 70B-shape MX-linear inference with the NCCL and the fused all-reduce.
N=${1:-2}
OUT=gpurun_out
export OMP_NUM_THREADS=4
run() { if [ "$N" = "1" ]; then python "$@"; else python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 200)) "$@"; fi; }
run bench.py --gpus $N --steps 20 --warmup 5 --skip-gemm 2> $OUT/r1_bench_n$N.err | tail -1 > $OUT/r1_bench_n$N.json
run tools/tp_llama_bench.py --mode quantize --model 70b 2> $OUT/tpq_n$N.err | tail -1 > $OUT/r1_tp_quantize_n$N.json
run tools/tp_llama_bench.py --mode infer --model 70b 2> $OUT/tpi_n$N.err | tail -1 > $OUT/r1_tp_infer_n$N.json
if [ "$N" != "1" ]; then run tools/tp_llama_bench.py --mode infer --model 70b --fused 2> $OUT/tpf_n$N.err | tail -1 > $OUT/r1_tp_infer_fused_n$N.json; fi
for f in $OUT/r1_bench_n$N.json $OUT/r1_tp_quantize_n$N.json $OUT/r1_tp_infer_n$N.json $OUT/r1_tp_infer_fused_n$N.json; do [ -f $f ] && (echo "== $f"; head -c 900 $f; echo); done
