mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,smsp__issue_active.avg.pct_of_peak_sustained_active,lts__t_bytes.sum,lts__t_bytes.sum.pct_of_peak_sustained_elapsed,l1tex__t_bytes.sum,l1tex__t_bytes.sum.pct_of_peak_sustained_elapsed
for c in 4 5; do
(timeout 900 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r4h_cfg$c.csv python tools/prof_cfg.py $c 3 3840x2160x1 > gpurun_out/r4h_cfg$c.log 2>&1)
(PTB_PROF_COUNT=1 timeout 300 python tools/prof_cfg.py $c 3 3840x2160x1 2>&1 | tail -2) > gpurun_out/r4h_cfg${c}_rays.log
R=$(grep RAYS gpurun_out/r4h_cfg${c}_rays.log | awk '{print $2}')
python tools/stage_metrics.py gpurun_out/r4h_cfg$c.csv 6531 $R > gpurun_out/r4h_cfg$c.md 2>&1
cat gpurun_out/r4h_cfg${c}_rays.log; cat gpurun_out/r4h_cfg$c.md
done
