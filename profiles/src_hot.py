"""Summarise an `ncu --page source --csv` dump: hottest SASS instructions by stall samples.
usage: python profiles/src_hot.py <source.csv> [top_n]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
data = rows[2:]
si, src, ie = hdr.index('# Samples'), hdr.index('Source'), hdr.index('Instructions Executed')
stall_cols = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
tot = sum(int(r[si]) for r in data if len(r) > si and r[si].isdigit())
print('total samples', tot, 'warp instructions', sum(int(r[ie]) for r in data if len(r) > ie and r[ie].isdigit()))
agg = {hdr[i]: sum(int(r[i]) for r in data if len(r) > i and r[i].isdigit()) for i in stall_cols}
print('stalls:', {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
top = sorted([(int(r[si]), i, r[src].strip(), r[ie]) for i, r in enumerate(data) if len(r) > si and r[si].isdigit()],
             reverse=True)[:n]
for t in sorted(top, key=lambda x: x[1]):
    r = data[t[1]]
    main = max(stall_cols, key=lambda i: int(r[i]) if r[i].isdigit() else 0)
    print('%6d  #%-5d %-60s exec=%s  %s' % (t[0], t[1], t[2][:60], t[3], hdr[main]))
