#!/bin/sh
# usage: tools/ncu_one.sh name <bench args...>   (one ncu --set full capture, summarised on the box)
name=$1; shift
python bench.py --steps 2 --warmup 1 --no-cpu --no-detail --no-density --no-e2e "$@" > gpurun_out/plain_$name.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"sample_kernel|density_kernel|ziggurat" -s 1 -c 1 -f \
    -o gpurun_out/prof_r02_$name python bench.py --steps 2 --warmup 1 --no-cpu --no-detail --no-density --no-e2e "$@" > gpurun_out/ncu_$name.log 2>&1
python tools/ncu_summary.py gpurun_out/prof_r02_$name.ncu-rep ${NTOP:-60} > gpurun_out/r02_${name}_ncu_summary.txt 2>&1
rm -f gpurun_out/prof_r02_$name.ncu-rep
