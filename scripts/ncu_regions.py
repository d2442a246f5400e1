#!/usr/bin/env python
"""Summarise an ncu source-page CSV: share of executed instructions / stall samples per SASS region and the hottest instructions.
usage: ncu -i rep --page source --csv --kernel-name regex:K > src.csv ; python scripts/ncu_regions.py src.csv [chunk] [launch-filter]"""
import csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
chunk = int(sys.argv[2]) if len(sys.argv) > 2 else 200
# the export may hold several launches back to back: split on header rows
blocks, cur = [], None
for r in rows:
    if r and r[0] == 'Kernel Name': cur = {'name': r[1], 'rows': []}; blocks.append(cur); continue
    if cur is None: continue
    if r and r[0] == 'Address': cur['H'] = r; continue
    cur['rows'].append(r)
for blk in blocks:
    H, R = blk['H'], blk['rows']
    iss, iex, isrc, itag = H.index('Warp Stall Sampling (All Samples)'), H.index('Instructions Executed'), H.index('Source'), H.index('L1 Tag Requests Global')
    tot = sum(int(r[iss] or 0) for r in R); totex = sum(int(r[iex] or 0) for r in R)
    if totex < 1e6: continue
    print('==', blk['name'], len(R), 'SASS; samples', tot, 'inst', totex, 'tags', sum(int(r[itag] or 0) for r in R))
    cols = ['stall_long_sb', 'stall_no_inst', 'stall_wait', 'stall_short_sb', 'stall_branch_resolving', 'stall_barrier', 'stall_selected', 'stall_not_selected', 'stall_math', 'stall_mio', 'stall_lg', 'stall_dispatch']
    ci = {c: H.index(c) for c in cols if c in H}
    for s in range(0, len(R), chunk):
        ch = R[s:s + chunk]
        sm = sum(int(r[iss] or 0) for r in ch); ex = sum(int(r[iex] or 0) for r in ch)
        d = {c[6:]: sum(int(r[i] or 0) for r in ch) for c, i in ci.items()}
        top = sorted(d.items(), key=lambda x: -x[1])[:3]
        ops = {}
        for r in ch:
            m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)', r[isrc])
            if m: ops[m.group(2)] = ops.get(m.group(2), 0) + 1
        print(f"{s:5d} samp {100*sm/tot:5.1f}% exec {100*ex/totex:5.1f}%", [(k, round(100 * v / max(sm, 1))) for k, v in top], sorted(ops.items(), key=lambda x: -x[1])[:4])
    for r in sorted(R, key=lambda r: -int(r[iss] or 0))[:12]:
        print(f"   {100*int(r[iss])/tot:5.2f}% ex={r[iex]:>9s} {r[isrc][:80]}")
