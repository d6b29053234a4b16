import csv, sys
rows=list(csv.reader(open(sys.argv[1])))
hdr=rows[0]; units=rows[1]
idx = int(sys.argv[2]) if len(sys.argv)>2 else -1
d=rows[2:][idx]
want = ['gpu__time_duration.sum','smsp__inst_executed.sum','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread',
 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','l1tex__data_pipe_lsu_wavefronts.sum',
 'l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed','smsp__issue_active.avg.pct_of_peak_sustained_active','dram__bytes_read.sum','dram__bytes_write.sum',
 'l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct','sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active','l1tex__m_xbar2l1tex_read_bytes.sum',
 'dram__throughput.avg.pct_of_peak_sustained_elapsed','lts__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum','l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum']
for h,u,v in zip(hdr,units,d):
    if h in want or ('issue_stalled' in h and h.endswith('_per_issue_active.ratio')) :
        try:
            f=float(v.replace(',',''))
        except: f=None
        if 'issue_stalled' in h and (f is None or f<0.3): continue
        print(f"{h:100s} {u:12s} {v}")
