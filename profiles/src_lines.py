"""Join an `ncu --page source --csv` SASS dump with `nvdisasm -g -c` output of the same cubin and
print executed warp instructions (and stall samples) per source line.
usage: python profiles/src_lines.py <source.csv> <nvdisasm.txt> <kernel-name-substring> [top_n]"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
data = [r for r in rows[2:] if len(r) > 10]
src, ie, si = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
key = sys.argv[3]
lines = open(sys.argv[2]).read().split('\n')
start = next(i for i, l in enumerate(lines) if l.startswith('.text.') and key in l)
ins, cur = [], None
for l in lines[start + 1:]:
    if l.startswith('//--------------------- .text') or l.startswith('.text.'):
        break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2)), 'inlined' in m.group(3))
        continue
    m = re.match(r'\s*/\*[0-9a-f]{4,}\*/\s+(.*?);', l)
    if m:
        ins.append((m.group(1).strip(), cur))
print(len(ins), 'SASS in cubin;', len(data), 'rows in csv')
by = collections.Counter()
bs = collections.Counter()
n = min(len(ins), len(data))
bad = 0
for k in range(n):
    a = data[k][src].split()
    b = ins[k][0].split()
    if a and b and a[0].split('.')[0] != b[0].split('.')[0] and not a[0].startswith('@'):
        bad += 1
    e = int(data[k][ie]) if data[k][ie].isdigit() else 0
    s = int(data[k][si]) if data[k][si].isdigit() else 0
    loc = ins[k][1][:2] if ins[k][1] else ('?', 0)
    by[loc] += e
    bs[loc] += s
print('opcode mismatches:', bad)
tot = sum(by.values())
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
for loc, e in by.most_common(top):
    print('%-18s:%-5d %9d %5.1f%%  samples %d' % (loc[0], loc[1], e, 100.0 * e / tot, bs[loc]))
# optional 5th argument: comma-separated "name:file:lo-hi" ranges to total (lines of inlined helpers count
# under their own file)
if len(sys.argv) > 5:
    for spec in sys.argv[5].split(','):
        name, f, rng = spec.split(':')
        lo, hi = map(int, rng.split('-'))
        e = sum(v for (ff, ln), v in by.items() if ff == f and lo <= ln <= hi)
        s = sum(v for (ff, ln), v in bs.items() if ff == f and lo <= ln <= hi)
        print('%-24s %9d %5.1f%%  samples %d' % (name, e, 100.0 * e / tot, s))
