#!/usr/bin/env python
"""Summarise an ncu report of the LK kernel: headline metrics, stall reasons, pipe utilisation and executed
warp-instructions per feature by source line / kernel phase (joins the SASS page with nvdisasm -g line info).
usage: ncu_report.py <rep.ncu-rep> <lib.so> <kernel mangled substring> <n_features> [top]"""
import csv, os, re, subprocess, sys, tempfile
from collections import defaultdict

rep, lib, kname, nfeat = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
tmp = tempfile.mkdtemp()
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
d = {h: (v, u) for h, u, v in zip(rows[0], rows[1], rows[2])}
keys = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed.avg.per_cycle_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__warps_eligible.avg.per_cycle_active', 'smsp__warps_active.avg.per_cycle_active',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum', 'lts__t_sectors_srcunit_tex.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'smsp__inst_executed.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__cycles_elapsed.avg', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_sector_hit_rate.pct']
for k in keys:
    if k in d:
        print('%-72s %18s %s' % (k, d[k][0], d[k][1]))
print('--- stalls per issue')
for h in rows[0]:
    if 'smsp__average_warps_issue_stalled' in h and h.endswith('_per_issue_active.ratio') and float(d[h][0] or 0) > 0.02:
        print('   %-40s %s' % (h.split('stalled_')[1].split('_per_issue')[0], d[h][0]))
print('--- pipes (% of peak, active)')
for h in rows[0]:
    if h.startswith('sm__inst_executed_pipe') and h.endswith('.avg.pct_of_peak_sustained_active') and float(d[h][0] or 0) > 0.5:
        print('   %-40s %s' % (h.split('pipe_')[1].split('.')[0], d[h][0]))
# line join
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
print('sass instrs: nvdisasm %d ncu %d' % (len(lines), len(ins)))
agg, ops = defaultdict(lambda: [0, 0]), defaultdict(int)
tot = ts = 0
for k in range(min(len(lines), len(ins))):
    ie = int(ins[k][ci['Instructions Executed']] or 0)
    st = int(ins[k][ci['Warp Stall Sampling (All Samples)']] or 0)
    agg[lines[k][0]][0] += ie; agg[lines[k][0]][1] += st
    tot += ie; ts += st
    t = ins[k][ci['Source']].split()
    op = (t[1] if t and t[0].startswith('@') else (t[0] if t else '?')).split('.')[0]
    ops[op] += ie
print('warp-instr per feature: %.1f' % (tot / nfeat))
srcfile = {}
def srcline(f, l):
    if f not in srcfile:
        p = os.path.join(os.path.dirname(os.path.abspath(lib)), '..', 'csrc', f)
        srcfile[f] = open(p).read().split('\n') if os.path.exists(p) else None
    return srcfile[f][l - 1].strip()[:100] if srcfile[f] and l - 1 < len(srcfile[f]) else ''
print('--- per source line: instr/feature, stall share')
for key, (ie, st) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    if key:
        print('%8.1f %5.1f%%  %s:%d  %s' % (ie / nfeat, 100.0 * st / max(ts, 1), key[0], key[1], srcline(*key)))
print('--- per source line by stall share')
for key, (ie, st) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:15]:
    if key:
        print('%8.1f %5.1f%%  %s:%d  %s' % (ie / nfeat, 100.0 * st / max(ts, 1), key[0], key[1], srcline(*key)))
print('--- opcodes (instr/feature)')
print('  '.join('%s %.0f' % (op, ie / nfeat) for op, ie in sorted(ops.items(), key=lambda kv: -kv[1])[:28]))
