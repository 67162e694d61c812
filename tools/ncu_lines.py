"""Per-CUDA-source-line warp-stall samples of an ncu report (needs -lineinfo + --import-source on):
    python tools/ncu_lines.py report.ncu-rep [top_n]
"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = next(r for r in rows if '# Samples' in r)
start = rows.index(hdr) + 1
ix = {h: i for i, h in enumerate(hdr)}
isamp = ix['# Samples']
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
agg, cur = {}, None
for r in rows[start:]:
    if len(r) != len(hdr):
        continue
    if r[0]:
        cur = (r[0], r[1])
        continue
    if r[2] == '...' or cur is None:
        continue
    a = agg.setdefault(cur, [0, 0, {}])
    a[0] += int(r[isamp] or 0)
    a[1] += int(r[ix['Instructions Executed']] or 0)
    for s in stalls:
        v = int(r[ix[s]] or 0)
        if v:
            a[2][s] = a[2].get(s, 0) + v
tot = sum(a[0] for a in agg.values())
print(f'total samples {tot}')
for (ln, src), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    st = ', '.join(f'{k[6:]} {v}' for k, v in sorted(a[2].items(), key=lambda kv: -kv[1])[:3])
    print(f'{100 * a[0] / tot:5.1f}%  inst {a[1]:>10d}  L{ln:>4s}  {src.strip()[:80]:80s} | {st}')
