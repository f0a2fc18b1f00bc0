#!/usr/bin/env python
"""Executed warp-instructions per feature by kernel phase, attributing inlined helpers to their outermost call site
(nvdisasm -gi).  usage: ncu_phases.py rep lib kernel_substr nfeat 'name:lo-hi,name:lo-hi,...' [source_file]"""
import csv, os, re, subprocess, sys, tempfile
from collections import defaultdict
rep, lib, kname, nfeat, spec = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4]), sys.argv[5]
srcfile = sys.argv[6] if len(sys.argv) > 6 else 'lk_fast.cu'
phases = []
for it in spec.split(','):
    n, r = it.split(':'); a, b = r.split('-'); phases.append((n, int(a), int(b)))
tmp = tempfile.mkdtemp()
subprocess.run('cd %s && cuobjdump -xelf all %s > /dev/null 2>&1' % (tmp, os.path.abspath(lib)), shell=True)
sass = ''
for fn in os.listdir(tmp):
    if fn.endswith('.cubin'):
        o = subprocess.run(['nvdisasm', '-gi', '-c', os.path.join(tmp, fn)], capture_output=True, text=True).stdout
        if kname in o:
            sass = o
lines, cur, inside = [], None, False
for ln in sass.splitlines():
    if ln.startswith('.text.'):
        inside = kname in ln
        continue
    if not inside:
        continue
    if '//## File' in ln:
        # outermost frame = last 'inlined at "file", line N' (or the line itself)
        ms = re.findall(r'"([^"]+)", line (\d+)', ln)
        body = [(f.split('/')[-1], int(l)) for f, l in ms]
        cur = (body[0], body[-1])
        continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', ln)
    if m:
        lines.append((cur, m.group(2).strip()))
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hdr = rows[1]; ci = {h: i for i, h in enumerate(hdr)}
ins = [r for r in rows[2:] if len(r) == len(hdr)]
assert len(ins) == len(lines), (len(ins), len(lines))
tot = defaultdict(float); ops = defaultdict(lambda: defaultdict(float)); stall = defaultdict(float); ts = 0
for k in range(len(ins)):
    ie = int(ins[k][ci['Instructions Executed']] or 0) / nfeat
    st = int(ins[k][ci['Warp Stall Sampling (All Samples)']] or 0)
    inner, outer = lines[k][0] if lines[k][0] else ((None, 0), (None, 0))
    ph = 'other'
    if outer[0] == srcfile:
        for n, a, b in phases:
            if a <= outer[1] < b:
                ph = n; break
    tot[ph] += ie; stall[ph] += st; ts += st
    t = ins[k][ci['Source']].split()
    op = (t[1] if t and t[0].startswith('@') else (t[0] if t else '?')).split('.')[0]
    ops[ph][op] += ie
print('%-28s %9s %7s   top opcodes' % ('phase', 'instr/ft', 'stall%'))
for n in [p[0] for p in phases] + ['other']:
    if tot[n] > 0:
        top = sorted(ops[n].items(), key=lambda kv: -kv[1])[:9]
        print('%-28s %9.1f %6.1f%%   %s' % (n, tot[n], 100 * stall[n] / max(ts, 1), ' '.join('%s:%.0f' % (o, v) for o, v in top)))
print('%-28s %9.1f' % ('TOTAL', sum(tot.values())))
if os.environ.get('DUMP_PHASE'):
    want = os.environ['DUMP_PHASE']
    for k in range(len(ins)):
        inner, outer = lines[k][0] if lines[k][0] else ((None, 0), (None, 0))
        ph = 'other'
        if outer[0] == srcfile:
            for n, a, b in phases:
                if a <= outer[1] < b:
                    ph = n; break
        if ph == want:
            ie = int(ins[k][ci['Instructions Executed']] or 0) / nfeat
            st = int(ins[k][ci['Warp Stall Sampling (All Samples)']] or 0)
            print('%5d %6.2f %6d  %-18s %-14s %s' % (k, ie, st, '%s:%d' % (inner[0][:12] if inner[0] else '?', inner[1]), 'at:%d' % outer[1], lines[k][1][:90]))
