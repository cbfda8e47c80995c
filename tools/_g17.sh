timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active -k regex:lg_nms --clock-control none --csv --log-file gpurun_out/t17_grid.csv python tools/stress_time.py 4 > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open("gpurun_out/t17_grid.csv")) if len(r)>10 and r[0].isdigit()]
for r in rows: print(r[0], r[4][:40], r[12], r[14])
PY
