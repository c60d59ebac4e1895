#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -n 3
python bench.py --steps 30 --warmup 5 > gpurun_out/bench_b.json 2> gpurun_out/bench_b.err; tail -n 5 gpurun_out/bench_b.err
python bench.py --steps 2 --warmup 3 --skip-e2e --skip-cpu > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'quantize|dequantize' -c 60 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --skip-e2e --skip-cpu > gpurun_out/ncu_launches.log 2>&1
python bench.py --steps 2 --warmup 3 --skip-e2e --skip-cpu > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'quantize_b32|dequantize_b32' -s 30 -c 4 -o gpurun_out/prof_qd python bench.py --steps 2 --warmup 3 --skip-e2e --skip-cpu > gpurun_out/ncu_full.log 2>&1
tail -n 3 gpurun_out/plain.log gpurun_out/ncu_launches.log gpurun_out/ncu_full.log
ls -la gpurun_out
