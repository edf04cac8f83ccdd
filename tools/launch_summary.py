"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: per-kernel time of the LAST tfr_process call."""
import csv, collections, sys
f = sys.argv[1]
rows = [r for r in csv.reader(open(f)) if len(r) > 10 and r[0].isdigit()]
names = [r[4].split('(')[0].replace('void ', '')[:44] for r in rows]
idx = [i for i, n in enumerate(names) if 'parse_kernel' in n]
a, b = (idx[-2] + 1, idx[-1] + 1) if len(idx) > 1 else (0, len(rows))
d = collections.OrderedDict()
for r, n in zip(rows[a:b], names[a:b]):
    if not (n.startswith('tfr::') or n.split('<')[0].endswith('_kernel')):   # the namespace is dropped when ncu ran with -k
        continue
    n = n if n.startswith('tfr::') else 'tfr::' + n
    d.setdefault(n, [0, 0.0])
    d[n][0] += 1
    d[n][1] += float(r[-1])
tot = sum(v[1] for v in d.values())
for n, (c, t) in d.items():
    print("  %10.1f us %4d  %s" % (t / 1e3, c, n))
print("  %10.1f us total (serialised, cold-cache per-launch times)" % (tot / 1e3))
