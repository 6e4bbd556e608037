#!/bin/bash
# round-2 final single-GPU batch (files r2z_*): the -m gpu suite on the product build and on the checked build, the bench lines of
# the three dense configs, the ncu launch list of the bench command, ncu --set full of the headline and config-3 launches,
# the per-warp trace.  Every ncu command is preceded by the same command exiting 0 without ncu.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2z_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2z_pytest.log
cp simplexmethod_b200/libenumgpu.so /tmp/libenumgpu_product.so
cp simplexmethod_b200/libenumgpu_check.so simplexmethod_b200/libenumgpu.so
( echo "# -m gpu suite against the checked build (make -C simplexmethod_b200/csrc check; -DENUMGPU_CHECK: bounds and alignment"; 
  echo "# asserts on every shared-window access of k_shared, queue and item-index asserts); deselected: the 21e9-basis m=12 n=44 run"; date -u;
  timeout 2400 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_parity.py::test_beyond_headline_size_properties 2>&1 ; echo "pytest rc=$?" ) > gpurun_out/r2z_check_build.log 2>&1
tail -3 gpurun_out/r2z_check_build.log
cp /tmp/libenumgpu_product.so simplexmethod_b200/libenumgpu.so
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/r2z_bench_12_40.json 2> gpurun_out/r2z_bench_12_40.err; echo "bench 12x40 rc=$?"
timeout 200 python bench.py --m 10 --n 30 --steps 20 --warmup 5 --cpu-ranks 30000000 > gpurun_out/r2z_bench_10_30.json 2> gpurun_out/r2z_bench_10_30.err; echo "bench 10x30 rc=$?"
timeout 200 python bench.py --m 8 --n 24 --steps 20 --warmup 5 > gpurun_out/r2z_bench_8_24.json 2> gpurun_out/r2z_bench_8_24.err; echo "bench 8x24 rc=$?"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2z_bench_ref.json 2> gpurun_out/r2z_bench_ref.err; echo "bench ref rc=$?"
# launch list of the bench command
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2z_plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2z_bench_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2z_ncu_bench.log 2>&1; echo "launch list rc=$?"
# full captures
python scripts/prof_run.py 12 40 0 0 2 2 > gpurun_out/r2z_plain_k12.log 2>&1 && \
ncu --set full --import-source on --clock-control none -k regex:k_shared -s 1 -c 1 -f -o gpurun_out/r2z_k_shared_m12n40 \
    python scripts/prof_run.py 12 40 0 0 2 2 > gpurun_out/r2z_ncu_k12.log 2>&1; echo "ncu 12x40 rc=$?"
python scripts/prof_run.py 10 30 0 0 2 2 > gpurun_out/r2z_plain_k10.log 2>&1 && \
ncu --set full --import-source on --clock-control none -k regex:k_shared -s 1 -c 1 -f -o gpurun_out/r2z_k_shared_m10n30 \
    python scripts/prof_run.py 10 30 0 0 2 2 > gpurun_out/r2z_ncu_k10.log 2>&1; echo "ncu 10x30 rc=$?"
for a in "12 40 0 0" "12 40 3 8" "10 30 0 0" "10 30 3 8" "8 24 0 0"; do echo "## $a"; python scripts/micro/trace_tail.py $a; done > gpurun_out/r2z_trace.log 2>&1
tail -4 gpurun_out/r2z_pytest.log
