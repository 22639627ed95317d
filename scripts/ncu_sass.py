"""SASS-level summary of an .ncu-rep source page: opcode mix (dynamic), stall samples by opcode and by stall reason."""
import csv, subprocess, sys, io, collections, re
rep = sys.argv[1]
src = subprocess.run(['ncu','-i',rep,'--page','source','--csv'],capture_output=True,text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = rows[1]
ix = {c:i for i,c in enumerate(h)}
mix = collections.Counter(); stall = collections.Counter(); reasons = collections.Counter()
tot_inst = 0; tot_samp = 0
recs = []
for r in rows[2:]:
    if len(r) < len(h): continue
    sass = r[ix['Source']].strip()
    m = re.match(r'(@!?U?P\d+\s+)?([A-Z0-9_.]+)', sass)
    op = m.group(2) if m else sass[:12]
    base = op.split('.')[0]
    if base in ('LDS','STS','LDG','STG'): base = op if op.count('.')<=2 else '.'.join(op.split('.')[:3])
    n = float(r[ix['Instructions Executed']] or 0); s = float(r[ix['Warp Stall Sampling (All Samples)']] or 0)
    mix[base] += n; stall[base] += s; tot_inst += n; tot_samp += s
    for c in h:
        if c.startswith('stall_') and 'Not Issued' not in c:
            reasons[c] += float(r[ix[c]] or 0)
    recs.append((s, n, sass))
print('total warp-instructions %.4g, samples %d' % (tot_inst, tot_samp))
print('-- dynamic opcode mix (top 25)')
for op,n in mix.most_common(25): print('  %-22s %6.2f%%  stall-samples %5.1f%%' % (op, 100*n/tot_inst, 100*stall[op]/max(tot_samp,1)))
print('-- stall reasons')
for c,v in reasons.most_common(10): print('  %-28s %5.1f%%' % (c, 100*v/max(sum(reasons.values()),1)))
print('-- hottest instructions')
for s,n,sass in sorted(recs, key=lambda x:-x[0])[:int(sys.argv[2]) if len(sys.argv)>2 else 20]:
    print('  %5.2f%%  n=%9d  %s' % (100*s/max(tot_samp,1), n, sass[:100]))
