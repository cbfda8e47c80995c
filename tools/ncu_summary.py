"""Summarise an .ncu-rep (raw page) into the handful of counters the roofline needs.
usage: ncu_summary.py <report.ncu-rep> [extra-metric-substring ...]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
extra = sys.argv[2:]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h = rows[0]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__block_size', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'launch__occupancy_limit_warps', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct', 'sm__cycles_elapsed.max',
        'sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32_sparsity_off.sum',
        'sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32_sparsity_off.sum.pct_of_peak_sustained_elapsed',
        'smsp__average_warp_latency_issue_stalled_long_scoreboard.pct', 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio']
for r in rows[2:]:
    print(r[h.index('ID')], r[h.index('Kernel Name')][:90])
    for k in h:
        if k in want or any(e in k for e in extra):
            print('   %-90s %14s %s' % (k, r[h.index(k)], rows[1][h.index(k)]))
