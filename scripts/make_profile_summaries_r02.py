"""Turns the captures of scripts/profile_r02.sh (gpurun_out/) into the tracked round-2 summaries under profiles/."""
import collections, csv, io, json, os, re, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.chdir(ROOT)


def sh(cmd):
    return subprocess.run(cmd, shell=True, capture_output=True, text=True).stdout


def copy(src, dst):
    if os.path.exists(src):
        shutil.copyfile(src, dst)
        print('copied', dst)
    else:
        print('MISSING', src)


# ---- plain copies ----
for name in ('r02_bench_n1.json', 'r02_bench_reference.json', 'r02_config_bench.txt', 'r02_pytest_gpu.log', 'r02_tensor_phase_trace.txt',
             'r02_small_kernel_trace.txt'):
    copy('gpurun_out/' + name, 'profiles/' + name)
if os.path.exists('gpurun_out/r02_small_batch.txt'):
    open('profiles/r02_small_batch.txt', 'w').write(
        '# scripts/small_probe.py: Adam step and residual-kernel time at the reference\'s batch sizes (CUDA events, no profiler)\n'
        '## small-batch kernel on (default: one 8-point batch per warp at most)\n' + open('gpurun_out/r02_small_batch.txt').read() +
        '## PINN_FUSED_SMALL_ROUNDS=0: 32-point kernel only (round 1)\n' + open('gpurun_out/r02_small_batch_off.txt').read())
if os.path.exists('gpurun_out/r02_admm_cancellation.txt'):
    open('profiles/r02_tanh_study.txt', 'w').write(
        '# tanh of the fused kernel: A/B of the variants (scripts/build_tanh_variants.sh) on B200\n'
        '## scripts/admm_cancellation_study.py: gradient error / |g| over 8 fresh ADMM-cancellation states (float64 oracle = truth)\n'
        + open('gpurun_out/r02_admm_cancellation.txt').read() +
        '## scripts/tanh_variants.py: the committed reference-run fixtures, oracle cases and the 16 Mi-point step time per variant\n'
        + open('gpurun_out/r02_tanh_variants.txt').read())
if os.path.exists('gpurun_out/r02_pytest_gpu.log'):
    lines = [l for l in open('gpurun_out/r02_pytest_gpu.log') if ' GPU [' in l or 'gradient error' in l or 'passed' in l or 'failed' in l]
    open('profiles/r02_converged_gpu.log', 'w').write(
        '# pytest tests -m gpu -s on B200: converged-accuracy lines (tests/test_converged_gpu.py) and the parity margins the tests print\n' + ''.join(lines))

# ---- launch list ----
if os.path.exists('gpurun_out/r02_launches.csv'):
    rows = list(csv.reader(open('gpurun_out/r02_launches.csv')))
    shutil.copyfile('gpurun_out/r02_launches.csv', 'profiles/r02_launches.csv')
    hi = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
    h = rows[hi]; iK, iV, iM, iU = h.index('Kernel Name'), h.index('Metric Value'), h.index('Metric Name'), h.index('Metric Unit')
    agg = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= iV or r[iM] != 'gpu__time_duration.sum':
            continue
        v = float(r[iV].replace(',', '')) * {'ns': 1e-3, 'us': 1.0, 'ms': 1e3}.get(r[iU], 1.0)
        a = agg.setdefault(r[iK].split('(')[0], [0, 0.0]); a[0] += 1; a[1] += v
    tot = sum(t for _, t in agg.values())
    out = ["# ncu launch list of: python bench.py --steps 2 --warmup 3 --nf-global 8388608 --nf-wide 262144 --cpu-points 65536  (B200, --metrics",
           "# gpu__time_duration.sum --clock-control none).  Per-launch times are cold-cache and serialised: compare SHARES.  The list covers the",
           "# whole bench line: timed region, e2e leg, sweep, the `extra` configs (tensor path, generic kernel, small-batch kernel) and the FFMA peak probe."]
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append("%-70s n=%4d total=%11.1f us  share=%5.1f%%  avg=%9.1f us" % (k[:70], n, t, 100 * t / tot, t / n))
    open('profiles/r02_launches.txt', 'w').write('\n'.join(out) + '\n')
    print('\n'.join(out[:14]))


# ---- ncu --set full summaries ----
def summarise(rep, title, srcfile, out, points=None, traffic_json=None, extra=''):
    if not os.path.exists(rep):
        print('MISSING', rep)
        return
    txt = "# %s (%s)\n" % (title, rep)
    txt += '\n'.join(l for l in sh('python scripts/ncu_summary.py %s 1' % rep).split('\n') if 'top source lines' not in l)
    txt += '\n'.join(sh('python scripts/ncu_sass.py %s 0' % rep).split('\n')[:45])
    txt += "\n== hottest source lines\n" + sh('python scripts/ncu_lines.py %s 30 %s' % (rep, srcfile))
    txt += extra
    open(out, 'w').write(txt)
    print('wrote', out)
    if points and traffic_json:
        raw = list(csv.reader(io.StringIO(sh('ncu -i %s --page raw --csv' % rep))))
        d = dict(zip(raw[0], raw[2])); un = dict(zip(raw[0], raw[1]))
        val = lambda k: float(d[k].replace(',', '')) * {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1}.get(un[k], 1)
        rd, wr = val('dram__bytes_read.sum'), val('dram__bytes_write.sum')
        json.dump({"kernel": title, "source": "ncu --set full --clock-control none, %s (summary: %s)" % (rep, out), "points_in_capture": points,
                   "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_per_point": round((rd + wr) / points, 1),
                   "algorithmic_hbm_bytes_per_point": 8}, open(traffic_json, 'w'), indent=1)


trace = open('gpurun_out/r02_tensor_phase_trace.txt').read() if os.path.exists('gpurun_out/r02_tensor_phase_trace.txt') else ''
summarise('gpurun_out/prof_tensor_r02.ncu-rep', 'ncu --set full --clock-control none of pinn_tc_kernel<4,1> (tcgen05 / TMEM / TMA path), [2,128x8,1], 75 776 points (4 tiles per CTA), B200',
          'pinns_b200/csrc/pinn_tensor.cu', 'profiles/r02_tensor_kernel_ncu.txt', 75776, 'profiles/r02_tensor_traffic.json',
          "\n== phase timeline of CTA 0 (scripts/tc_phase_trace.py, clock64, -DPINN_TC_TRACE build)\n" + trace)
summarise('gpurun_out/prof_fused_r02.ncu-rep', 'ncu --set full --clock-control none of pinn_fused_kernel<20,true>, 2 Mi points, B200',
          'pinns_b200/csrc/pinn_fused.cu', 'profiles/r02_fused_kernel_ncu.txt', 2097152, 'profiles/r02_fused_traffic.json')
summarise('gpurun_out/prof_small_r02.ncu-rep', 'ncu --set full --clock-control none of pinn_fused_small_kernel<20,true>, 1000 points + 100 data points, B200',
          'pinns_b200/csrc/pinn_fused.cu', 'profiles/r02_small_kernel_ncu.txt')

# ---- static SASS opcode histograms of every kernel of the library (cuobjdump, no GPU needed) ----
out = ["# cuobjdump -sass of pinns_b200/csrc/*.o (sm_100a): static opcode histogram per kernel.  Blackwell-native evidence: UTCHMMA = tcgen05.mma,",
       "# LDTM / STTM = tcgen05.ld / st, UTMALDG = cp.async.bulk.tensor (TMA tensor copy), UBLKCP = cp.async.bulk (TMA engine), SYNCS = mbarrier, UTCBAR = tcgen05.commit, FFMA2 = fma.rn.f32x2, ACQBULK / PREEXIT = griddepcontrol.wait / launch_dependents."]
for obj in sorted(os.listdir('pinns_b200/csrc')):
    if not obj.endswith('.o'):
        continue
    sass = sh('cuobjdump -sass pinns_b200/csrc/%s' % obj)
    cur, hist = None, collections.OrderedDict()
    for ln in sass.split('\n'):
        m = re.search(r'Function : (\S+)', ln)
        if m:
            cur = sh('c++filt %s' % m.group(1)).strip().replace('(anonymous namespace)::', '').replace('<unnamed>::', '')
            cur = re.sub(r'^void\s+', '', cur).split('(')[0]
            hist[cur] = collections.Counter()
            continue
        m = re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+(.*?);', ln)
        if cur and m:
            p = m.group(1).split()
            op = p[1] if p[0].startswith('@') and len(p) > 1 else p[0]
            hist[cur][op.split('.')[0]] += 1
    for k, c in hist.items():
        key = ['UTCHMMA', 'LDTM', 'STTM', 'UTMALDG', 'UBLKCP', 'UTCBAR', 'SYNCS', 'FFMA2', 'FFMA', 'LDS', 'STS', 'LDG', 'STG', 'LDGSTS', 'RED', 'MUFU', 'SHFL', 'CCTL', 'BAR', 'ACQBULK', 'PREEXIT']
        out.append("%-14s %-58s total %5d | %s" % (obj, k[:58], sum(c.values()), ' '.join('%s:%d' % (o, c[o]) for o in key if c.get(o))))
open('profiles/r02_sass_opcodes.txt', 'w').write('\n'.join(out) + '\n')
print('\n'.join(out))
