"""Per-source-line and per-opcode executed-instruction counts of an ncu report (needs -lineinfo + --import-source on):
    python tools/ncu_inst.py report.ncu-rep [top_n]
"""
import collections
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
lines, ops, cur = {}, collections.Counter(), None
for r in rows[start:]:
    if len(r) != len(hdr):
        continue
    if r[0]:
        cur = (r[0], r[1])
        continue
    if r[2] == '...' or cur is None:
        continue
    n = int(r[ix['Instructions Executed']] or 0)
    lines[cur] = lines.get(cur, 0) + n
    t = r[3].split()
    op = t[1] if t[0].startswith('@') else t[0]
    ops['.'.join(op.split('.')[:2])] += n
tot = sum(lines.values())
print(f'{tot} warp instructions')
for (ln, src), n in sorted(lines.items(), key=lambda kv: -kv[1])[:top]:
    print(f'{100 * n / tot:5.1f}% {n:>10d} L{ln:>4s} {src.strip()[:110]}')
print('opcodes:')
for k, v in ops.most_common(top):
    print(f'{100 * v / tot:5.1f}% {v:>10d} {k}')
