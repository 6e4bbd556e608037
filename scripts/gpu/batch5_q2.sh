#!/bin/bash
# two-stage survivor stacks (no call in the d loop): parity on the product and the checked build, then old vs new kernel times
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2g_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2g_pytest.log
tail -3 gpurun_out/r2g_pytest.log
cp simplexmethod_b200/libenumgpu.so /tmp/libenumgpu_product.so
cp simplexmethod_b200/libenumgpu_check.so simplexmethod_b200/libenumgpu.so
( echo "# -m gpu suite against the checked build (make -C simplexmethod_b200/csrc check; -DENUMGPU_CHECK: bounds and alignment";
  echo "# asserts on every shared-window access of k_shared, stack and item-index asserts); deselected: the 21e9-basis m=12 n=44 run"; date -u;
  timeout 1200 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_parity.py::test_beyond_headline_size_properties 2>&1 ; echo "pytest rc=$?" ) > gpurun_out/r2g_check_build.log 2>&1
tail -3 gpurun_out/r2g_check_build.log
cp /tmp/libenumgpu_product.so simplexmethod_b200/libenumgpu.so
timeout 600 python scripts/gpu/kbench.py scripts/gpu/variants/head.so scripts/gpu/variants/q2full.so scripts/gpu/variants/head.so scripts/gpu/variants/q2full.so 2>&1 | tee gpurun_out/r2g_kbench.log
