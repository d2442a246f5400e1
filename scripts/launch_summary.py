#!/usr/bin/env python
"""Aggregate an ncu `--metrics gpu__time_duration.sum --csv` launch list per kernel."""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1], errors='ignore')))
hdr = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
H = rows[hdr]; kn = H.index('Kernel Name'); mv = H.index('Metric Value')
agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
for r in rows[hdr + 1:]:
    if len(r) <= mv: continue
    try: v = float(r[mv].replace(',', ''))
    except ValueError: continue
    n = r[kn].split('(')[0]; agg[n][0] += 1; agg[n][1] += v; agg[n][2] = max(agg[n][2], v)
tot = sum(v[1] for v in agg.values())
for n, (c, t, mx) in sorted(agg.items(), key=lambda x: -x[1][1])[:10]:
    print(f"{n[:48]:48s} {c:5d} {t/1e6:9.3f} ms {100*t/tot:5.1f}%  avg {t/c/1e3:8.1f} us  max {mx/1e3:8.1f} us")
