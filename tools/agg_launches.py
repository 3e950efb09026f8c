# dev tool: aggregate an ncu --csv launch list by kernel for the LAST sampling step (between the last two ddpm_step launches)
import csv, collections, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
h = rows[hdr]; ki = h.index('Kernel Name'); mi = h.index('Metric Name'); vi = h.index('Metric Value'); ui = h.index('Metric Unit'); idi = h.index('ID')
per = {}
for r in rows[hdr + 1:]:
    if len(r) <= vi: continue
    d = per.setdefault(int(r[idi]), {'name': re.sub(r'\(.*', '', r[ki])})
    v = float(r[vi].replace(',', '')); u = r[ui]
    v *= {'us': 1e3, 'usecond': 1e3, 'ms': 1e6, 'msecond': 1e6, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(u, 1)
    d[r[mi]] = v
ids = sorted(per)
dd = [i for i in ids if per[i]['name'].startswith('ddpm_step')]
lo, hi = dd[-2] + 1, dd[-1]
agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
for i in ids:
    if lo <= i <= hi:
        d = per[i]; a = agg[d['name']]
        a[0] += 1; a[1] += d['gpu__time_duration.sum'] / 1e3; a[2] += d.get('dram__bytes_read.sum', 0) + d.get('dram__bytes_write.sum', 0); a[3] += d.get('lts__t_bytes.sum', 0)
tot = sum(a[1] for a in agg.values())
print(f'one step: {tot:.1f} us over {sum(a[0] for a in agg.values())} launches')
for n, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{n[:50]:50s} n={a[0]:3d} {a[1]:9.1f} us {100*a[1]/tot:5.1f}% {a[1]/a[0]:8.2f} us/launch dram {a[2]/1e6:8.1f} MB L2 {a[3]/1e6:9.1f} MB")
t = [(per[i]['gpu__time_duration.sum'] / 1e3, per[i].get('lts__t_bytes.sum', 0) / 1e6) for i in ids if lo <= i <= hi and per[i]['name'].startswith('tc5v2')]
print('tc5v2 us/L2MB:', ' '.join(f"{a:.1f}/{b:.0f}" for a, b in t[:12]))
