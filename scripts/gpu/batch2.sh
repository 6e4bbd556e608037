#!/bin/bash
# round-2 GPU batch 2: -m gpu suite on the product build and on the checked build (-DENUMGPU_CHECK), trace of the small configs
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2h_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2h_pytest.log
tail -3 gpurun_out/r2h_pytest.log
cp simplexmethod_b200/libenumgpu.so /tmp/libenumgpu_product.so
cp simplexmethod_b200/libenumgpu_check.so simplexmethod_b200/libenumgpu.so
( echo "# -m gpu suite against the checked build (make -C simplexmethod_b200/csrc check; -DENUMGPU_CHECK)"; date -u;
  timeout 2400 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_parity.py::test_beyond_headline_size_properties 2>&1 ; echo "pytest rc=$?" ) > gpurun_out/r2h_check_build.log 2>&1
tail -4 gpurun_out/r2h_check_build.log
cp /tmp/libenumgpu_product.so simplexmethod_b200/libenumgpu.so
for a in "10 30 0 0" "8 24 0 0" "12 40 3 8"; do echo "## $a"; python scripts/micro/trace_tail.py $a; done > gpurun_out/r2h_trace.log 2>&1
cat gpurun_out/r2h_trace.log | grep "record written"
