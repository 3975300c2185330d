#!/bin/sh
# one ncu --set full capture per compute-bound sampler (development aid; run under gpurun)
set -x
run() {  # name dist streams draws extra...
  name=$1; shift
  python bench.py --steps 2 --warmup 1 --no-cpu --no-detail --no-density --no-e2e "$@" > gpurun_out/plain_$name.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:"sample_kernel|density_kernel" -s 1 -c 1 -f \
      -o gpurun_out/prof_r01_$name python bench.py --steps 2 --warmup 1 --no-cpu --no-detail --no-density --no-e2e "$@" > gpurun_out/ncu_$name.log 2>&1
  # the reports are too big to bring back together (64 MiB limit): summarise on the box
  python tools/ncu_summary.py gpurun_out/prof_r01_$name.ncu-rep 30 > gpurun_out/r01_${name}_ncu_summary.txt 2>&1
  rm -f gpurun_out/prof_r01_$name.ncu-rep
}
run exponential --dist exponential_rate --streams 262144 --draws 1000
run box_muller --dist normal_box_muller --streams 262144 --draws 1000
run binomial_d1 --dist binomial --streams 4194304 --draws 1
run poisson_d1 --dist poisson --streams 4194304 --draws 1
run binomial_d16 --dist binomial --streams 1048576 --draws 16
run gamma --dist gamma_scale --streams 1048576 --draws 16
run hypergeometric --dist hypergeometric --streams 524288 --draws 16
run density_poisson --workload density --density poisson --obs 50000000
ls -la gpurun_out/
