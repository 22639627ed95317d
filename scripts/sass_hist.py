"""Static opcode histogram of one kernel in an object file (cuobjdump -sass), split at backward branches."""
import subprocess, sys, collections, re
obj, pat = sys.argv[1], sys.argv[2]
out = subprocess.run(['cuobjdump', '-sass', obj], capture_output=True, text=True).stdout
cur, ins = None, []
for ln in out.split('\n'):
    m = re.search(r'Function : (\S+)', ln)
    if m: cur = m.group(1); continue
    if cur and pat in cur:
        m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', ln)
        if m: ins.append((int(m.group(1), 16), m.group(2).strip()))
def opname(t):
    p = t.split()
    op = p[1] if p[0].startswith('@') else p[0]
    return op.split('.')[0] + ('.128' if '.128' in op else '')
print('total static instructions', len(ins))
print(' '.join('%s:%d' % kv for kv in collections.Counter(opname(t) for _, t in ins).most_common(24)))
# loops: backward BRA targets
for a, t in ins:
    m = re.search(r'BRA\s+(?:\S+\s+)?`?\(?\.?L_x_\d+\)?|BRA.*0x([0-9a-f]+)', t)
    if 'BRA' in t:
        m = re.search(r'0x([0-9a-f]+)', t)
        if m and int(m.group(1), 16) < a:
            lo = int(m.group(1), 16)
            body = [x for x in ins if lo <= x[0] <= a]
            c = collections.Counter(opname(x[1]) for x in body)
            if len(body) > 60:
                print('loop %05x..%05x n=%d | %s' % (lo, a, len(body), ' '.join('%s:%d' % kv for kv in c.most_common(10))))
