#!/usr/bin/env python
"""Join an ncu SASS source page (ncu -i X.ncu-rep --page source --csv --print-source sass) with nvdisasm -g line info
to get executed instructions / stall samples per CUDA source line.  usage: ncu_lines.py lk_src.csv lkfast.sass <kernel substring> [top]"""
import csv, re, sys
from collections import defaultdict

def main():
    src_csv, sass, kname = sys.argv[1], sys.argv[2], sys.argv[3]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    # nvdisasm: find the .text section of the kernel, collect (line, opcode) in order
    lines, cur, inside = [], None, False
    for ln in open(sass):
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
    rows = list(csv.reader(open(src_csv)))
    hdr = rows[1]
    ci = {h: i for i, h in enumerate(hdr)}
    ins = [r for r in rows[2:] if len(r) == len(hdr)]
    print('nvdisasm instrs', len(lines), 'ncu instrs', len(ins))
    n = min(len(lines), len(ins))
    agg = defaultdict(lambda: [0, 0, 0])
    tot_i = tot_s = 0
    opagg = defaultdict(int)
    for k in range(n):
        r = ins[k]
        ie = int(r[ci['Instructions Executed']] or 0)
        st = int(r[ci['Warp Stall Sampling (All Samples)']] or 0)
        key = lines[k][0]
        agg[key][0] += ie; agg[key][1] += st; agg[key][2] += 1
        tot_i += ie; tot_s += st
        op = r[ci['Source']].split()[0] if r[ci['Source']].split() else '?'
        if op.startswith('@'):
            op = r[ci['Source']].split()[1]
        opagg[op.split('.')[0]] += ie
    print('total warp-instr', tot_i, 'stall samples', tot_s)
    print('--- by source line (instr%, stall%, #sass)')
    for key, (ie, st, cnt) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print('%5.1f%% %5.1f%% %4d  %s' % (100.0 * ie / tot_i, 100.0 * st / max(tot_s, 1), cnt, key))
    print('--- by opcode')
    for op, ie in sorted(opagg.items(), key=lambda kv: -kv[1])[:25]:
        print('%5.1f%%  %s' % (100.0 * ie / tot_i, op))

main()
