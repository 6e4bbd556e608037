#!/bin/bash
# ncu --set full of the headline launch for each given variant name (scripts/gpu/variants/<name>.so)
for v in "$@"; do
  python scripts/gpu/krun.py scripts/gpu/variants/$v.so 12 40 2 || exit 1
  ncu --set full --import-source on --clock-control none -k regex:k_shared -s 1 -c 1 -f -o gpurun_out/r2_$v \
      python scripts/gpu/krun.py scripts/gpu/variants/$v.so 12 40 2 > gpurun_out/r2_$v.ncu.log 2>&1
  tail -2 gpurun_out/r2_$v.ncu.log
done
