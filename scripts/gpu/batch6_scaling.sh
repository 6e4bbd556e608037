#!/bin/bash
# round-2 closing scaling batch (gpurun --gpus 8), final kernel: headline at 8 / 4 / 2 GPUs, config 3 at 8 (most
# important first: the call's time limit is whatever is left of the round's budget).  N = 1 comes from the 1-GPU batch.
mkdir -p gpurun_out
run() {  # name N extra-args...
  local name=$1 n=$2; shift 2
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) \
      bench.py --gpus $n "$@" > gpurun_out/r2y_${name}_n$n.json 2> gpurun_out/r2y_${name}_n$n.err
  echo "$name N=$n rc=$? $(head -c 200 gpurun_out/r2y_${name}_n$n.json)"
}
run m12n40 8 --steps 10 --warmup 3 --no-cpu-baseline
run m12n40 4 --steps 10 --warmup 3 --no-cpu-baseline
run m12n40 2 --steps 10 --warmup 3 --no-cpu-baseline
run m10n30 8 --lp-m 10 --lp-n 30 --steps 50 --warmup 5 --no-cpu-baseline
