"""SASS-address-ordered profile of an .ncu-rep: consecutive instructions grouped into blocks split at branches /
labels, with stall samples, executed warp-instructions and the opcode mix -- shows which LOOP the time is in when
every FFMA2 shares one inline-asm source line."""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if 'Warp Stall Sampling (All Samples)' in r)
h = rows[hi]
iA, iS, iC, iN = h.index('Address'), h.index('Source'), h.index('Warp Stall Sampling (All Samples)'), h.index('Instructions Executed')
ins = []
for r in rows[hi + 1:]:
    if len(r) < len(h): continue
    try: ins.append((r[iA], r[iS].strip(), float(r[iC] or 0), float(r[iN] or 0)))
    except ValueError: pass
tot_s = sum(i[2] for i in ins); tot_n = sum(i[3] for i in ins)
# split into blocks where the executed count changes by more than 2x (loop nest boundaries) or at branches
blocks, cur = [], []
for k, it in enumerate(ins):
    if cur and (it[3] > 1.5 * cur[-1][3] or it[3] < cur[-1][3] / 1.5):
        blocks.append(cur); cur = []
    cur.append(it)
    if it[1].split()[0].startswith('BRA') or (it[1].startswith('@') and 'BRA' in it[1]):
        blocks.append(cur); cur = []
if cur: blocks.append(cur)
print('total samples %d  warp-instr %.4g  blocks %d' % (tot_s, tot_n, len(blocks)))
for b in blocks:
    s = sum(i[2] for i in b); n = sum(i[3] for i in b)
    if s < 0.004 * tot_s and n < 0.004 * tot_n: continue
    ops = collections.Counter()
    for i in b:
        t = i[1].split()
        op = t[1] if t[0].startswith('@') and len(t) > 1 else t[0]
        ops[op.split('.')[0] + ('.128' if '.128' in op else '')] += 1
    print('%s..%s  n_instr=%4d  exec/instr=%.3g  samples %5.1f%%  instr %5.1f%%  samples/instr-exec=%.2f | %s' % (
        b[0][0][-5:], b[-1][0][-5:], len(b), n / len(b), 100 * s / tot_s, 100 * n / tot_n, (s / tot_s) / max(n / tot_n, 1e-9),
        ' '.join('%s:%d' % kv for kv in ops.most_common(7))))
