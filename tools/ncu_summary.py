import csv,sys,subprocess,collections
rep=sys.argv[1]; nruns=float(sys.argv[2]) if len(sys.argv)>2 else 100000
raw=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(raw.splitlines())); hdr=rows[0]; units=rows[1]; r=rows[2]
def g(k):
    return r[hdr.index(k)] if k in hdr else 'n/a'
keys=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','dram__throughput.avg.pct_of_peak_sustained_elapsed','smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__grid_size','launch__block_size','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum']
for k in keys: print('%-70s %s %s'%(k,g(k),units[hdr.index(k)] if k in hdr else ''))
print('inst/run', float(g('smsp__inst_executed.sum'))/nruns)
for k in hdr:
    if k.startswith('smsp__average_warps_issue_stalled') and k.endswith('per_issue_active.ratio'):
        v=float(r[hdr.index(k)])
        if v>0.05: print('  stall %-40s %.2f'%(k.replace('smsp__average_warps_issue_stalled_','').replace('_per_issue_active.ratio',''),v))
