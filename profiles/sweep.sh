#!/bin/bash
# usage: profiles/sweep.sh out_name size steps variant...   (run on the GPU box)
out=gpurun_out/$1; size=$2; steps=$3; shift 3
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader > $out
for v in "$@"; do
  if [ "$v" = "default" ]; then lib=anuga_core_b200/libswk.so; else lib=variants/libswk_$v.so; fi
  SWK_LIB=$PWD/$lib timeout 300 python profiles/kernel_bench.py $size $steps 2>&1 | tail -1 >> $out
done
cat $out
