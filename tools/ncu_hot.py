"""Per-kernel-instance SASS profile from an ncu source-page CSV dump: instruction totals and the hottest instructions.
usage: ncu -i rep --page source --csv --kernel-name regex:NAME --print-source sass > src.csv ; python tools/ncu_hot.py src.csv [instance] [top]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
inst = int(sys.argv[2]) if len(sys.argv) > 2 else -1
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
kern = []
cur = None
for r in rows:
    if r and r[0] == 'Kernel Name':
        cur = {'name': r[1], 'rows': [], 'hdr': None}; kern.append(cur); continue
    if r and r[0] == 'Address':
        cur['hdr'] = r; continue
    if cur is not None and cur['hdr'] is not None and len(r) >= len(cur['hdr']) - 1:
        cur['rows'].append(r)
for i, k in enumerate(kern):
    h = k['hdr']; iI = h.index('Instructions Executed'); iS = h.index('# Samples')
    tot = sum(int(r[iI]) for r in k['rows']); ts = sum(int(r[iS]) for r in k['rows'])
    print(i, k['name'][:70], 'warp-instr', tot, 'samples', ts)
if inst >= 0:
    k = kern[inst]; h = k['hdr']
    iI = h.index('Instructions Executed'); iS = h.index('# Samples'); iSrc = h.index('Source'); iT = h.index('Avg. Threads Executed')
    tot = sum(int(r[iI]) for r in k['rows'])
    ops = collections.Counter(); opsamp = collections.Counter()
    for r in k['rows']:
        t = r[iSrc].split()
        op = t[1] if t[0].startswith('@') else t[0]
        ops[op.split('.')[0]] += int(r[iI]); opsamp[op.split('.')[0]] += int(r[iS])
    print('opcode executed share:', [(o, round(100 * c / tot, 1)) for o, c in ops.most_common(25)])
    print('opcode stall-sample share:', [(o, c) for o, c in opsamp.most_common(15)])
    print('--- instructions in address order with exec >= 0.2% of total')
    for r in k['rows']:
        if int(r[iI]) >= 0.002 * tot:
            print(f"{int(r[iI]):>10d} {int(r[iS]):>6d} {r[iT]:>5s}  {r[iSrc].strip()[:90]}")
