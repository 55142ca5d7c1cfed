"""Prints a compact table from an `ncu --page raw --csv` export."""
import csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
want = [('gpu__time_duration.sum','ms'),('dram__bytes_read.sum','rdMB'),('dram__bytes_write.sum','wrMB'),
 ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','dram%'),('sm__throughput.avg.pct_of_peak_sustained_elapsed','sm%'),
 ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active','tensor%'),('sm__warps_active.avg.pct_of_peak_sustained_active','occ%'),
 ('launch__registers_per_thread','regs'),('launch__grid_size','grid'),('l1tex__throughput.avg.pct_of_peak_sustained_elapsed','l1%'),
 ('lts__throughput.avg.pct_of_peak_sustained_elapsed','l2%'),('smsp__inst_executed.sum','winst')]
kn = hdr.index('Kernel Name')
print('kernel'.ljust(30), ' '.join(n.rjust(9) for _, n in want))
for r in data:
    m = re.search(r'(\w+_kernel)(<[^>]*>)?', r[kn]); name = (m.group(1)[:18] + (m.group(2) or ''))[:30]
    vals = []
    for w, n in want:
        if w not in hdr: vals.append('-'); continue
        i = hdr.index(w); v = float(r[i].replace(',', '') or 0); u = units[i]
        if n in ('rdMB','wrMB'):
            v = v * {'Gbyte':1e3,'Mbyte':1,'Kbyte':1e-3,'byte':1e-6}.get(u,1)
        vals.append(f"{v:.3f}" if v < 100 else f"{v:.0f}")
    print(name.ljust(30), ' '.join(v.rjust(9) for v in vals))
