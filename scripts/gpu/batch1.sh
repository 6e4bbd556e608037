#!/bin/bash
# round-2 GPU batch 1: full -m gpu suite, headline bench, configs 2 and 3 at N=1
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv > gpurun_out/r2a_smi.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
tail -5 gpurun_out/r2a_pytest.log
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/r2a_bench_12_40.json 2> gpurun_out/r2a_bench_12_40.err; echo "bench rc=$?"
timeout 200 python bench.py --m 10 --n 30 --steps 20 --warmup 5 --cpu-ranks 30000000 > gpurun_out/r2a_bench_10_30.json 2> gpurun_out/r2a_bench_10_30.err; echo "bench rc=$?"
timeout 200 python bench.py --m 8 --n 24 --steps 20 --warmup 5 > gpurun_out/r2a_bench_8_24.json 2> gpurun_out/r2a_bench_8_24.err; echo "bench rc=$?"
head -c 600 gpurun_out/r2a_bench_12_40.json; echo
