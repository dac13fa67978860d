"""Per-source-line opcode breakdown of an `ncu --page source --csv --print-source cuda,sass` dump."""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = None
file = None
per_line = collections.defaultdict(collections.Counter)
samples = collections.Counter()
line_src = {}
cur = None
for r in rows:
    if not r:
        continue
    if r[0] == 'File Path':
        file = r[1].split('/')[-1]
        continue
    if r[0] == 'Function Name':
        continue
    if r[0] == 'Line No':
        hdr = r
        continue
    if hdr is None:
        continue
    if r[0] != '':
        cur = (file, r[0])
        line_src[cur] = r[1]
    elif cur and len(r) > 7 and r[2] not in ('-', '...'):
        try:
            n = int(r[7])
        except ValueError:
            continue
        m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)', r[3])
        per_line[cur][m.group(2) if m else '?'] += n
        try:
            samples[cur] += int(r[6])
        except ValueError:
            pass
tot = sum(sum(c.values()) for c in per_line.values())
ops = collections.Counter()
for c in per_line.values():
    ops.update(c)
print('total warp-instructions', tot)
print('by opcode:', ', '.join('%s %.1f%%' % (k, 100.0 * v / tot) for k, v in ops.most_common(14)))
for (f, ln), c in sorted(per_line.items(), key=lambda kv: -sum(kv[1].values()))[:top]:
    n = sum(c.values())
    print("%5.1f%% samp=%-6d %s:%s  %s\n        %s" % (100.0 * n / tot, samples[(f, ln)], f, ln, line_src[(f, ln)].strip()[:100], dict(c.most_common(5))))
