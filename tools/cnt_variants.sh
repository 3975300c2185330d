M=smsp__inst_executed.sum,smsp__thread_inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active
cp monty_b200/libmonty_b200.so /tmp/keep.so
for v in base rf32 rf8; do
  cp variants/$v/libmonty_b200.so monty_b200/libmonty_b200.so
  python tools/count_run.py truncated_normal binomial > /dev/null 2>&1
  ncu --metrics $M --clock-control none -k regex:"sorted_sample" --csv --log-file gpurun_out/cnt_$v.csv python tools/count_run.py truncated_normal binomial > /dev/null 2>&1
done
cp /tmp/keep.so monty_b200/libmonty_b200.so
