"""Per-CUDA-source-line stall samples and executed instructions from an .ncu-rep (needs -lineinfo)."""
import csv, subprocess, sys, io, collections
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
srcfile = sys.argv[3] if len(sys.argv) > 3 else 'pinns_b200/csrc/pinn_fused.cu'
out = subprocess.run(['ncu','-i',rep,'--page','source','--csv','--print-source','cuda,sass'],capture_output=True,text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i,r in enumerate(rows) if 'Warp Stall Sampling (All Samples)' in r)
h = rows[hi]
iS = h.index('Warp Stall Sampling (All Samples)'); iN = h.index('Instructions Executed')
stall_cols = [(i,c) for i,c in enumerate(h) if c.startswith('stall_') and 'Not Issued' not in c]
lines = {}
for r in rows[hi+1:]:
    if len(r) < len(h) or not r[0].strip().isdigit(): continue
    try:
        ln = int(r[0]); s = float(r[iS] or 0); n = float(r[iN] or 0)
    except ValueError:
        continue
    try:
        st = {c: float(r[i] or 0) for i,c in stall_cols}
    except ValueError:
        continue
    if ln in lines:
        lines[ln][0] += s; lines[ln][1] += n
        for c in st: lines[ln][2][c] += st[c]
    else:
        lines[ln] = [s, n, st]
src = open(srcfile).read().split('\n')
tot = sum(v[0] for v in lines.values()); totn = sum(v[1] for v in lines.values())
print('total samples %d, warp-instructions %.4g' % (tot, totn))
for ln,(s,n,st) in sorted(lines.items(), key=lambda kv:-kv[1][0])[:top]:
    top2 = sorted(st.items(), key=lambda kv:-kv[1])[:3]
    print('%5.1f%% samp %5.1f%% inst  L%-4d %-70s | %s' % (100*s/tot, 100*n/totn, ln, src[ln-1].strip()[:70], ' '.join('%s=%d%%' % (c.replace('stall_',''), 100*v/max(s,1)) for c,v in top2)))
