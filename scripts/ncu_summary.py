"""Summarise an .ncu-rep: headline metrics + stall reasons + per-source-line hot spots."""
import csv, subprocess, sys, collections, io
rep = sys.argv[1]
raw = subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
for vals in rows[2:]:
    d = dict(zip(hdr, vals))
    print('== kernel', d.get('Kernel Name','?')[:60], 'grid', d.get('Grid Size'), 'block', d.get('Block Size'))
    for k in ['gpu__time_duration.sum','launch__registers_per_thread','smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active',
              'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active','sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
              'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed','l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum','l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum',
              'dram__bytes_read.sum','dram__bytes_write.sum','lts__t_bytes.sum','sm__cycles_active.avg','sm__cycles_elapsed.avg','smsp__warps_active.avg.per_cycle_active','sm__throughput.avg.pct_of_peak_sustained_elapsed']:
        if k in d: print('  %-75s %s %s' % (k, d[k], units[hdr.index(k)]))
    st = {k:float(v) for k,v in d.items() if k.startswith('smsp__average_warps_issue_stalled') and k.endswith('_per_issue_active.ratio')}
    for k,v in sorted(st.items(), key=lambda kv:-kv[1])[:8]:
        print('  stall %-40s %.3f' % (k.replace('smsp__average_warps_issue_stalled_','').replace('_per_issue_active.ratio',''), v))
src = subprocess.run(['ncu','-i',rep,'--page','source','--csv','--print-source','cuda'],capture_output=True,text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
if rows:
    h = rows[0]
    def col(name):
        for i,c in enumerate(h):
            if c.strip()==name: return i
        return None
    ci = {n:col(n) for n in ['#','Source','Warp Stall Sampling (All Samples)','Instructions Executed','Warp Stall Sampling (Not-issued Samples)']}
    tot = 0; lines=[]
    for r in rows[1:]:
        try:
            s = float(r[ci['Warp Stall Sampling (All Samples)']] or 0); n = float(r[ci['Instructions Executed']] or 0)
        except Exception: continue
        tot += s; lines.append((s, n, r[ci['#']], r[ci['Source']][:110]))
    print('== top source lines by stall samples (total %d)' % tot)
    for s,n,ln,txt in sorted(lines, key=lambda x:-x[0])[:int(sys.argv[2]) if len(sys.argv)>2 else 25]:
        print('  %5.1f%% inst=%10d  L%-4s %s' % (100*s/max(tot,1), n, ln, txt.strip()))
