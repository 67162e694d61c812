"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel (used for profiles/*.md)."""
import collections
import csv
import re
import sys


def main(path, top=30):
    lines = [l for l in open(path) if not l.startswith('==')]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        if row.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        v = float(row['Metric Value'].replace(',', ''))
        unit = row['Metric Unit']
        v = v / 1e6 if unit.startswith('n') else v / 1e3 if unit.startswith('u') else v
        short = re.sub(r'\(.*', '', row['Kernel Name'])[:90]
        agg[short][0] += 1
        agg[short][1] += v
    tot = sum(v[1] for v in agg.values())
    n = sum(v[0] for v in agg.values())
    print(f'{n} launches, {tot:.1f} ms of kernel time (cold-cache, serialised: compare SHARES)')
    print(f'{"ms":>10} {"share":>6} {"n":>6}  kernel')
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f'{v[1]:10.2f} {100 * v[1] / tot:5.1f}% {v[0]:6d}  {k}')


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30)
