#!/usr/bin/env python
"""Dump SASS of a kernel from an ncu report with per-instruction executed counts per feature and source line.
usage: ncu_sass.py rep lib kernel_substr nfeat [min_per_feat] [max_per_feat]"""
import csv, os, re, subprocess, sys, tempfile
rep, lib, kname, nfeat = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
lo = float(sys.argv[5]) if len(sys.argv) > 5 else 0
hi = float(sys.argv[6]) if len(sys.argv) > 6 else 1e9
tmp = tempfile.mkdtemp()
subprocess.run('cd %s && cuobjdump -xelf all %s > /dev/null 2>&1' % (tmp, os.path.abspath(lib)), shell=True)
sass = ''
for fn in os.listdir(tmp):
    if fn.endswith('.cubin'):
        o = subprocess.run(['nvdisasm', '-g', '-c', os.path.join(tmp, fn)], capture_output=True, text=True).stdout
        if kname in o:
            sass = o
lines, cur, inside = [], None, False
for ln in sass.splitlines():
    if ln.startswith('.text.'):
        inside = kname in ln
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2)))
        continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', ln)
    if m:
        lines.append((cur, m.group(2).strip()))
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hdr = rows[1]
ci = {h: i for i, h in enumerate(hdr)}
ins = [r for r in rows[2:] if len(r) == len(hdr)]
for k in range(min(len(lines), len(ins))):
    ie = int(ins[k][ci['Instructions Executed']] or 0) / nfeat
    st = int(ins[k][ci['Warp Stall Sampling (All Samples)']] or 0)
    if lo <= ie <= hi:
        l = lines[k][0]
        print('%5d %7.2f %6d  %-22s %s' % (k, ie, st, '%s:%d' % (l[0][:14], l[1]) if l else '?', lines[k][1][:100]))
