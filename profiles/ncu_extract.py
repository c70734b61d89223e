import csv, sys, subprocess
rep=sys.argv[1]
out=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
hdr=rows[0]; units=rows[1]
keys=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','sm__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__throughput.avg.pct_of_peak_sustained_active','lts__throughput.avg.pct_of_peak_sustained_elapsed','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__occupancy_limit_shared_mem','launch__occupancy_limit_registers','launch__shared_mem_per_block_dynamic','launch__grid_size','launch__waves_per_multiprocessor','sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','lts__t_sector_hit_rate.pct','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','smsp__inst_executed.sum','sm__cycles_elapsed.max','smsp__cycles_active.avg','sm__cycles_active.avg','lts__t_bytes.sum','l1tex__t_bytes_pipe_lsu_mem_global_op_ld.sum','l1tex__t_sector_hit_rate.pct']
keys+= [h for h in hdr if 'issue_stalled' in h and 'per_issue_active' in h]
for k in keys:
    if k in hdr:
        i=hdr.index(k); print(f"{k[:100]:100s} {units[i]:10s}", [r[i] for r in rows[2:]])
