import csv,sys,subprocess
rep=sys.argv[1]
out=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
hdr=rows[0]; units=rows[1]
want=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','dram__throughput.avg.pct_of_peak_sustained_elapsed','sm__throughput.avg.pct_of_peak_sustained_elapsed','lts__t_sector_hit_rate.pct','l1tex__t_sector_hit_rate.pct','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','launch__grid_size','l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum','l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum','l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum','l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum','l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum','l1tex__t_requests_pipe_lsu_mem_global_op_st.sum','smsp__thread_inst_executed_per_inst_executed.ratio','lts__t_sectors_op_write.sum','lts__t_sectors_op_read.sum','sm__inst_executed_pipe_fp64.sum']
for r in rows[2:]:
    print('----', r[hdr.index('Kernel Name')])
    for w in want:
        if w in hdr: print('  %-70s %s %s'%(w, r[hdr.index(w)], units[hdr.index(w)]))
    items=[]
    for i,h in enumerate(hdr):
        if h.startswith('smsp__pcsamp_warps_issue_stalled') and not h.endswith('not_issued'):
            try: items.append((float(r[i]),h.replace('smsp__pcsamp_warps_issue_stalled_','')))
            except: pass
    tot=sum(v for v,_ in items) or 1
    print('  stalls:', ', '.join('%s %.0f%%'%(h,100*v/tot) for v,h in sorted(items,reverse=True)[:7]))
