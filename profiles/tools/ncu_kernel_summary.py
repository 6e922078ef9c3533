# usage: python ncu_kernel_summary.py report.ncu-rep  -> per launch: time, issue/occupancy, DRAM bytes, shared-memory wavefronts, stall ratios, busiest pipes
import csv, subprocess, sys, io
rep=sys.argv[1]
raw=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(io.StringIO(raw)))
hdr=rows[0]
def g(r,k):
    return r[hdr.index(k)] if k in hdr else 'n/a'
want=['gpu__time_duration.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__warps_active.avg.pct_of_peak_sustained_active','dram__bytes_read.sum','dram__bytes_write.sum','launch__registers_per_thread','smsp__inst_executed.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','launch__waves_per_multiprocessor','dram__throughput.avg.pct_of_peak_sustained_elapsed','lts__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__throughput.avg.pct_of_peak_sustained_elapsed']
for r in rows[2:]:
    print('----', g(r,'Kernel Name')[:60], 'grid', g(r,'launch__grid_size'), 'block', g(r,'launch__block_size'))
    for w in want: print('  ',w, g(r,w))
    st=[]
    for i,h in enumerate(hdr):
        if h.startswith('smsp__average_warps_issue_stalled_') and h.endswith('_per_issue_active.ratio'):
            try: st.append((float(r[i]),h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]))
            except: pass
    print('   stalls:', ', '.join(f'{n} {v:.2f}' for v,n in sorted(st,reverse=True)[:7]))
    pipes=[]
    for i,h in enumerate(hdr):
        if h.startswith('sm__inst_executed_pipe_') and h.endswith('.avg.pct_of_peak_sustained_active'):
            try:
                v=float(r[i])
                if v>5: pipes.append((v,h[len('sm__inst_executed_pipe_'):].split('.')[0]))
            except: pass
    print('   pipes:', ', '.join(f'{n} {v:.0f}' for v,n in sorted(set(pipes),reverse=True)))
