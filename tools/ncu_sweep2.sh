#!/bin/sh
set -x
run() {
  name=$1; shift
  python bench.py --steps 2 --warmup 1 --no-cpu --no-detail --no-density --no-e2e "$@" > gpurun_out/plain_$name.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:"sample_kernel|density_kernel" -s 1 -c 1 -f \
      -o gpurun_out/prof_r01_$name python bench.py --steps 2 --warmup 1 --no-cpu --no-detail --no-density --no-e2e "$@" > gpurun_out/ncu_$name.log 2>&1
  python tools/ncu_summary.py gpurun_out/prof_r01_$name.ncu-rep 40 > gpurun_out/r01_${name}_ncu_summary.txt 2>&1
  rm -f gpurun_out/prof_r01_$name.ncu-rep
}
run sorted_binomial_d16 --dist binomial --streams 1048576 --draws 16
run sorted_hypergeometric --dist hypergeometric --streams 524288 --draws 16
run sorted_poisson_d1 --dist poisson --streams 4194304 --draws 1
run density_poisson_table --workload density --density poisson --obs 50000000
