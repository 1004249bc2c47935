"""One line per kernel launch of an .ncu-rep: duration and the counters used in profiles/README.md."""
import csv, subprocess, sys
rep = sys.argv[1]
want = [('gpu__time_duration.sum', 'ms'), ('dram__bytes_read.sum', 'rd'), ('dram__bytes_write.sum', 'wr'),
        ('smsp__thread_inst_executed_per_inst_executed.ratio', 'thr/inst'), ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue%'),
        ('l1tex__t_sector_hit_rate.pct', 'L1hit%'), ('lts__t_sector_hit_rate.pct', 'L2hit%'), ('sm__inst_executed.sum', 'winst'),
        ('dram__throughput.avg.pct_of_peak_sustained_elapsed', 'dram%'), ('sm__warps_active.avg.pct_of_peak_sustained_active', 'occ%'),
        ('launch__registers_per_thread', 'regs')]
out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
print('kernel'.ljust(34), ' '.join(f"{n:>10s}" for _, n in want))
for r in rows[2:]:
    name = r[hdr.index('Kernel Name')].split('(')[0].replace('<unnamed>::', '').replace('void ', '')[:34]
    vals = []
    for w, n in want:
        if w in hdr:
            v = r[hdr.index(w)].replace(',', '')
            u = units[hdr.index(w)]
            try:
                f = float(v)
                if n == 'ms': f *= {'ns': 1e-6, 'us': 1e-3, 'ms': 1, 's': 1e3}.get(u, 1)
                if n in ('rd', 'wr'): f *= {'byte': 1e-9, 'Kbyte': 1e-6, 'Mbyte': 1e-3, 'Gbyte': 1}.get(u, 1)
                vals.append(f"{f:10.3f}" if f < 1e6 else f"{f:10.3e}")
            except ValueError:
                vals.append(f"{v:>10s}")
        else:
            vals.append(' ' * 10)
    print(name.ljust(34), ' '.join(vals))
