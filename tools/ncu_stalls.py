#!/usr/bin/env python
"""Summarise an `ncu --page source --csv` dump: total stall samples by reason, and the hottest SASS lines.
usage: ncu_stalls.py src.csv [top_n [kernel-substring]]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
want = sys.argv[3] if len(sys.argv) > 3 else None   # kernel-name substring: only that kernel's first block
# the dump holds one block per captured launch ("Kernel Name" row, header row, one row per SASS line)
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
pick = next((i for i in starts if want is None or want in rows[i][1]), starts[0])
end = next((i for i in starts if i > pick), len(rows))
print("kernel:", rows[pick][1])
rows = rows[pick:end]
hdr = rows[1]
idx = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith('stall_') and '(Not Issued)' not in h]
tot = collections.Counter()
lines = []
for r in rows[2:]:
    if len(r) < len(hdr) or r[0] in ("Address", "Kernel Name"): continue
    n = int(r[idx['# Samples']] or 0)
    for c in stall_cols:
        tot[c] += int(r[idx[c]] or 0)
    lines.append((n, r[idx['Source']].strip(), int(r[idx['Instructions Executed']] or 0),
                  {c: int(r[idx[c]] or 0) for c in stall_cols if int(r[idx[c]] or 0)}))
all_s = sum(tot.values())
print('total samples', all_s)
for c, v in tot.most_common(12):
    print('  %-26s %8d  %5.1f%%' % (c, v, 100.0 * v / max(all_s, 1)))
ops = collections.Counter()
for n, s, ie, d in lines:
    ops[s.split()[0] if not s.startswith('@') else s.split()[1]] += ie
print('instructions executed by opcode:')
for k, v in ops.most_common(14):
    print('  %-14s %10d' % (k, v))
print('hottest lines:')
for n, s, ie, d in sorted(lines, key=lambda x: -x[0])[:top]:
    print('  %6d %9d  %-60s %s' % (n, ie, s[:60], dict(sorted(d.items(), key=lambda kv: -kv[1])[:3])))
