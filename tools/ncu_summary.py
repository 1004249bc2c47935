"""Summarise an .ncu-rep: per-kernel key metrics (+ optional hot-loop SASS of one kernel)."""
import csv, subprocess, sys, collections
rep = sys.argv[1]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__grid_size', 'launch__registers_per_thread',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'l1tex__throughput.avg.pct_of_peak_sustained_active', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__inst_executed.sum', 'smsp__warps_eligible.avg.per_cycle_active']
out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    name = r[hdr.index('Kernel Name')].split('(')[0][-40:]
    print('---', name)
    for w in want:
        if w in hdr:
            print(f"    {w:70s} {r[hdr.index(w)]:>16s} {units[hdr.index(w)]}")
if len(sys.argv) > 2:
    kname = sys.argv[2]
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', 'regex:' + kname], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    kern = []; cur = None
    for r in rows:
        if r and r[0] == 'Kernel Name':
            cur = {'name': r[1], 'rows': [], 'hdr': None}; kern.append(cur); continue
        if r and r[0] == 'Address':
            cur['hdr'] = r; continue
        if cur is not None and cur['hdr'] is not None and len(r) == len(cur['hdr']):
            cur['rows'].append(r)
    k = max(kern, key=lambda k: sum(int(r[k['hdr'].index('Instructions Executed')]) for r in k['rows']))
    h = k['hdr']; iI = h.index('Instructions Executed'); iS = h.index('Source'); iSm = h.index('# Samples')
    tot = sum(int(r[iI]) for r in k['rows']); totS = sum(int(r[iSm]) for r in k['rows'])
    mx = max(int(r[iI]) for r in k['rows'])
    hot = [r for r in k['rows'] if int(r[iI]) > 0.7 * mx]
    print(f"kernel {k['name'][:60]} total warp-instr {tot} samples {totS}; hot-loop SASS lines {len(hot)} (exec ~{mx})")
    ops = collections.Counter()
    for r in hot:
        t = r[iS].split()
        op = t[1] if t[0].startswith('@') else t[0]
        ops[op.split('.')[0]] += 1
    print("   hot loop opcode counts:", dict(ops.most_common(30)))
    if len(sys.argv) > 3:
        for r in hot: print(r[iI], r[iSm], r[iS].strip()[:100])
