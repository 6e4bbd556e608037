#!/bin/bash
# round-2 scaling batch (gpurun --gpus 8): headline (m=12, n=40) and config 3 (m=10, n=30) at 1/2/4/8 GPUs, config 2
# (m=8, n=24) at 1, and the in-process multi-device test on distinct devices.
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.max.sm --format=csv > gpurun_out/r2s_smi.txt 2>&1
run() {  # name N extra-args...
  local name=$1 n=$2; shift 2
  if [ "$n" = 1 ]; then
    timeout 300 python bench.py --gpus 1 "$@" > gpurun_out/r2s_${name}_n1.json 2> gpurun_out/r2s_${name}_n1.err
  else
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) \
        bench.py --gpus $n "$@" > gpurun_out/r2s_${name}_n$n.json 2> gpurun_out/r2s_${name}_n$n.err
  fi
  echo "$name N=$n rc=$? $(head -c 160 gpurun_out/r2s_${name}_n$n.json)"
}
for n in 1 2 4 8; do run m12n40 $n --steps 10 --warmup 3 --no-cpu-baseline; done
for n in 1 2 4 8; do run m10n30 $n --lp-m 10 --lp-n 30 --steps 50 --warmup 5 --no-cpu-baseline; done
run m8n24 1 --lp-m 8 --lp-n 24 --steps 50 --warmup 5 --no-cpu-baseline
timeout 600 python -m pytest tests/test_gpu_parity.py -v -k "multi_device" > gpurun_out/r2s_multi_device_pytest.log 2>&1; echo "multi-device pytest rc=$?" | tee -a gpurun_out/r2s_multi_device_pytest.log
tail -3 gpurun_out/r2s_multi_device_pytest.log
