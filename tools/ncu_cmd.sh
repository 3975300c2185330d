#!/bin/sh
# usage: tools/ncu_cmd.sh name kernel_regex skip <command...>   (one `ncu --set full` capture of the
# (skip+1)-th launch matching kernel_regex, summarised on the box; the command is first run plain)
name=$1; regex=$2; skip=$3; shift 3
"$@" > gpurun_out/plain_$name.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"$regex" -s $skip -c 1 -f \
    -o gpurun_out/prof_r02_$name "$@" > gpurun_out/ncu_$name.log 2>&1
python tools/ncu_summary.py gpurun_out/prof_r02_$name.ncu-rep ${NTOP:-60} > gpurun_out/r02_${name}_ncu_summary.txt 2>&1
rm -f gpurun_out/prof_r02_$name.ncu-rep
