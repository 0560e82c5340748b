"""Join an `ncu --page source --csv` dump (dynamic instruction counts per SASS instruction) with `nvdisasm --print-line-info`
of the same kernel (source file:line per SASS instruction): executed thread instructions and stall samples per source line.
usage: sass_line_profile.py <source.csv> <kernel.sass> [n_symbols]"""
import csv, collections, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; data = rows[2:]
nsym = float(sys.argv[3]) if len(sys.argv) > 3 else 3.24e6
iS = hdr.index('Source'); iE = hdr.index('Instructions Executed'); iSm = hdr.index('# Samples')
cur = None; loc = []
stack = ''
for l in open(sys.argv[2]):
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2)))
        continue
    if re.match(r'\s*/\*([0-9a-f]+)\*/\s+', l):
        loc.append(cur)
assert len(loc) == len(data), (len(loc), len(data))
ex = collections.Counter(); sm = collections.Counter(); ops = collections.defaultdict(collections.Counter)
tot = 0; totS = 0
for r, c in zip(data, loc):
    e = int(r[iE]); s = int(r[iSm])
    ex[c] += e; sm[c] += s; tot += e; totS += s
    parts = r[iS].split(); o = parts[1] if parts[0].startswith('@') else parts[0]
    ops[c][o.split('.')[0]] += e
print('warp inst', tot, 'thread-inst/symbol', tot * 32 / nsym)
for c, e in ex.most_common(int(sys.argv[4]) if len(sys.argv) > 4 else 60):
    print(f'{c[0]:22s}:{c[1]:4d} {e / tot * 100:6.2f}% thr/sym={e * 32 / nsym:7.0f} samples {sm[c] / totS * 100:5.1f}%  ',
          {k: round(v * 32 / nsym) for k, v in ops[c].most_common(5)})
