#!/bin/sh
# Instruction counts and DRAM traffic per launch of every sampler kernel of the bench's tables
# (one light ncu pass, under gpurun): writes gpurun_out/counts.csv + count_manifest.json;
# tools/make_inst_counts.py turns them into profiles/inst_counts.json and profiles/traffic.json.
M=smsp__inst_executed.sum,smsp__thread_inst_executed.sum,sm__inst_executed_pipe_fp64.sum,sm__inst_executed_pipe_alu.sum,sm__inst_executed_pipe_fma.sum,sm__inst_executed_pipe_xu.sum,sm__inst_executed_pipe_lsu.sum,gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
python tools/count_run.py "$@" > gpurun_out/count_plain.log 2>&1 &&
ncu --metrics $M --clock-control none -k regex:"sample_kernel|ziggurat_kernel" --csv --log-file gpurun_out/counts.csv \
    python tools/count_run.py "$@" > gpurun_out/count_ncu.log 2>&1
