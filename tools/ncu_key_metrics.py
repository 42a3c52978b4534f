import csv,sys,subprocess
rep=sys.argv[1]
out=subprocess.run(["ncu","-i",rep,"--page","raw","--csv"],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
hdr=rows[0]; vals=rows[-1]
want=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','launch__registers_per_thread','sm__warps_active.avg.pct_of_peak_sustained_active','sm__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','smsp__thread_inst_executed_per_inst_executed.ratio','smsp__inst_executed.sum','launch__grid_size','smsp__issue_active.avg.pct_of_peak_sustained_active','l1tex__throughput.avg.pct_of_peak_sustained_elapsed','lts__throughput.avg.pct_of_peak_sustained_elapsed','launch__occupancy_limit_registers','launch__waves_per_multiprocessor','sass__inst_executed_local_loads','sass__inst_executed_local_stores','sm__inst_executed_pipe_xu.sum','smsp__inst_executed_pipe_xu.sum','sm__inst_executed_pipe_fma.sum','sm__inst_executed_pipe_alu.sum']
print("kernel:",vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "")
for i,h in enumerate(hdr):
    if h in want: print(" ",h, rows[1][i], vals[i])
    if 'stall' in h and 'ratio' in h:
        try:
            if float(vals[i] or 0)>0.25: print("  STALL",h.replace("smsp__average_warps_issue_stalled_","").replace("_per_issue_active.ratio",""), vals[i])
        except: pass
