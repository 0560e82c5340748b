"""Summarise an `ncu --page source --csv` dump: instruction mix, stall samples, hot regions."""
import csv, collections, sys
path = sys.argv[1]; nsym = float(sys.argv[2]) if len(sys.argv) > 2 else 3.24e6
rows = list(csv.reader(open(path)))
hdr = rows[1]; data = rows[2:]
iS = hdr.index('Source'); iE = hdr.index('Instructions Executed'); iSm = hdr.index('# Samples')
stall_cols = [i for i, h in enumerate(hdr) if h.startswith('stall_') or h.startswith('Stall')]
tot = sum(int(r[iE]) for r in data); totS = sum(int(r[iSm]) for r in data)
op = collections.Counter(); ops = collections.Counter()
for r in data:
    parts = r[iS].split()
    o = parts[1] if parts[0].startswith('@') else parts[0]
    o = o.split('.')[0]
    op[o] += int(r[iE]); ops[o] += int(r[iSm])
print('total warp inst', tot, 'thread-inst/symbol', tot * 32 / nsym, 'samples', totS)
for o, c in op.most_common(28):
    print(f'{o:10s} {c / tot * 100:6.2f}%  thr-inst/sym={c * 32 / nsym:8.0f}  samples {ops[o] / totS * 100:5.1f}%')
# hot regions: split at BAR instructions
print('--- regions between barriers (by SASS order) ---')
reg = []; cur = [0, 0, 0, None]
for k, r in enumerate(data):
    cur[0] += int(r[iE]); cur[1] += int(r[iSm]); cur[2] += 1
    if cur[3] is None: cur[3] = k
    if 'BAR.SYNC' in r[iS]:
        reg.append(cur); cur = [0, 0, 0, None]
reg.append(cur)
for c in reg:
    print(f'sass[{c[3]}..+{c[2]}] inst {c[0] / tot * 100:5.1f}%  samples {c[1] / totS * 100:5.1f}%')
