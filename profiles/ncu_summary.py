import csv, sys, subprocess, re, collections
rep=sys.argv[1]
raw=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(raw.splitlines()))
hdr=rows[0]; units=rows[1]; data=rows[2:]
want=[('Kernel Name','name'),('gpu__time_duration.sum','us'),('dram__bytes_read.sum','rdMB'),('dram__bytes_write.sum','wrMB'),
 ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','dram%'),('sm__warps_active.avg.pct_of_peak_sustained_active','occ%'),
 ('launch__registers_per_thread','regs'),('smsp__issue_active.avg.pct_of_peak_sustained_active','issue%'),
 ('sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active','fma%'),('sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active','alu%'),
 ('l1tex__throughput.avg.pct_of_peak_sustained_active','l1%'),('l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','bankconf'),
 ('l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','smwave'),('smsp__inst_executed.sum','inst'),
 ('smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio','st_long'),
 ('smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio','st_short'),
 ('smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio','st_mio'),
 ('smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio','st_bar'),
 ('smsp__average_warps_issue_stalled_wait_per_issue_active.ratio','st_wait'),
 ('smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio','st_math'),
 ('smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio','st_notsel'),
 ('smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio','st_lg'),
 ('lts__t_sector_hit_rate.pct','l2hit%')]
idx=[(hdr.index(w),n) for w,n in want if w in hdr]
seen=set()
for d in data:
    name=re.sub(r'\(.*','',d[hdr.index('Kernel Name')]).replace('void ','').replace('fft::','')
    if name in seen: continue
    seen.add(name)
    out=[]
    for i,n in idx:
        v=d[i]
        if n=='name': out.append(name); continue
        try:
            f=float(v.replace(',',''))
            if n in('rdMB','wrMB'): 
                u=units[i]; f = f/1e6 if u=='byte' else (f if u=='Mbyte' else f*1e3 if u=='Gbyte' else f/1e3)
            out.append(f"{n}={f:.3g}")
        except: out.append(f"{n}={v}")
    print(' '.join(out))
