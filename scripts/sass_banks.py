"""Register-bank analysis of FFMA instructions in a kernel's SASS (even/odd banks, B300_MICROARCH.md 'RF banking').
usage: sass_banks.py file.o kernel-substring"""
import re, subprocess, sys, collections
obj, pat = sys.argv[1], sys.argv[2]
txt = subprocess.run(['cuobjdump','-sass',obj],capture_output=True,text=True).stdout
funcs = re.split(r'\n\s*Function : ', txt)
for f in funcs[1:]:
    name = f.split('\n',1)[0]
    if pat not in name: continue
    tot = conf = three = 0
    prev_src = None
    ex = []
    for m in re.finditer(r'FFMA(?:\.\w+)* (R\d+), (-?R\d+)(\.reuse)?, (-?R\d+|-?c\[[^\]]+\]\[[^\]]+\]|-?[0-9.e+-]+|-?UR\d+)(\.reuse)?, (-?R\d+)(\.reuse)?', f):
        tot += 1
        regs = [m.group(2), m.group(4), m.group(6)]
        srcs = []
        for slot, r in enumerate(regs):
            r = r.lstrip('-')
            if not r.startswith('R'): continue
            # operand served by the reuse cache if the previous instruction had .reuse on the same slot & register
            cached = prev_src is not None and prev_src[slot] == (r, True)
            if not cached: srcs.append(int(r[1:]))
        ev = len(set(x for x in srcs if x % 2 == 0)); od = len(set(x for x in srcs if x % 2 == 1))
        if max(ev, od) >= 2:
            conf += 1
            if len(ex) < 6: ex.append(m.group(0))
        prev_src = [(regs[0].lstrip('-'), bool(m.group(3))), (regs[1].lstrip('-'), bool(m.group(5))), (regs[2].lstrip('-'), bool(m.group(7)))]
    print(name[:90]); print('  FFMA %d, with >=2 fresh operands in one bank: %d (%.1f%%)' % (tot, conf, 100.0*conf/max(tot,1)))
    for e in ex: print('   ', e)
