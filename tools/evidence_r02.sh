#!/bin/sh
# Round-2 evidence pass (under gpurun): bench line, ncu summaries of the main kernels, instruction
# counts, launch list.  Outputs under gpurun_out/; the summaries are then copied to profiles/.
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err
NTOP=70 sh tools/ncu_one.sh ziggurat_2p24x128
NTOP=70 sh tools/ncu_one.sh ziggurat_c2 --streams 1048576 --draws 1000
NTOP=60 sh tools/ncu_one.sh box_muller --dist normal_box_muller --streams 1048576 --draws 1000
NTOP=60 sh tools/ncu_one.sh exponential --dist exponential_rate --streams 1048576 --draws 1000
NTOP=60 sh tools/ncu_one.sh hypergeometric_d16 --dist hypergeometric --streams 2097152 --draws 16
NTOP=60 sh tools/ncu_one.sh binomial_d1 --dist binomial --streams 4194304 --draws 1
NTOP=60 sh tools/ncu_one.sh gamma_d16 --dist gamma_scale --streams 2097152 --draws 16
NTOP=60 sh tools/ncu_cmd.sh density_poisson density_tile 1 python tools/density_run.py poisson 250000000 2
NTOP=60 sh tools/ncu_cmd.sh density_nbinom "density_sorted|density_kernel" 1 python tools/density_run.py negative_binomial_mu 100000000 2
sh tools/ncu_counts.sh
python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/launch_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/launch_ncu.log 2>&1
ls -la gpurun_out | tail -30
