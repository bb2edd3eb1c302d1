"""Group an `ncu --page source --csv` SASS dump into straight-line regions with equal execution
counts and print the regions that execute the most warp instructions.
usage: python profiles/src_regions.py <source.csv> [top_n]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
data = [r for r in rows[2:] if len(r) > 10]
src, ie, si = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
first = []
for i, r in enumerate(data):
    if i > 0 and r[src].strip() == data[0][src].strip() and data[i + 1][src].strip() == data[1][src].strip():
        break
    first.append(r)
tot = sum(int(r[ie]) for r in first if r[ie].isdigit())
print(len(first), 'SASS instructions; warp instructions executed:', tot)
byop = collections.Counter()
for r in first:
    if r[ie].isdigit():
        t = r[src].split()
        op = t[1] if t[0].startswith('@') else t[0]
        byop[op.split('.')[0]] += int(r[ie])
print(byop.most_common(16))
reg, cur = [], None
for i, r in enumerate(first):
    e = int(r[ie]) if r[ie].isdigit() else 0
    s = int(r[si]) if r[si].isdigit() else 0
    if cur and cur[2] == e:
        cur[1] = i
        cur[3] += s
    else:
        cur = [i, i, e, s]
        reg.append(cur)
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
big = sorted(reg, key=lambda x: -(x[1] - x[0] + 1) * x[2])[:n]
for a, b, e, s in sorted(big):
    print('#%d-%d len %d exec %d total %d (%.1f%%) samples %d   %s' % (a, b, b - a + 1, e, (b - a + 1) * e,
          100.0 * (b - a + 1) * e / tot, s, first[a][src].strip()[:50]))
