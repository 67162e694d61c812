"""Summarise an `ncu --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum] --csv` launch list by
kernel (used for profiles/*.txt).   python tools/summarize_launches.py launches.csv [top] [first_launch_id] [end_launch_id]
first_launch_id / end_launch_id: keep launches with first <= ncu ID < end (one forward out of several)."""
import collections
import csv
import re
import sys


def to_ms(v, unit):
    return v / 1e6 if unit.startswith('n') else v / 1e3 if unit.startswith('u') else v * 1e3 if unit == 's' else v


def to_mb(v, unit):
    u = unit.lower()
    return v / 1e6 if u == 'byte' else v / 1e3 if u.startswith('k') else v * 1e3 if u.startswith('g') else v


def main(path, top=30, first_id=0, end_id=1 << 60):
    lines = [l for l in open(path) if not l.startswith('==')]
    per = collections.defaultdict(dict)           # launch id -> {name, ms, rd, wr}
    for row in csv.DictReader(lines):
        i = int(row['ID'])
        if i < first_id or i >= end_id:
            continue
        v = float(row['Metric Value'].replace(',', ''))
        d = per[i]
        d['name'] = re.sub(r'\(.*', '', row['Kernel Name'])[:90]
        if row['Metric Name'] == 'gpu__time_duration.sum':
            d['ms'] = to_ms(v, row['Metric Unit'])
        elif row['Metric Name'] == 'dram__bytes_read.sum':
            d['rd'] = to_mb(v, row['Metric Unit'])
        elif row['Metric Name'] == 'dram__bytes_write.sum':
            d['wr'] = to_mb(v, row['Metric Unit'])
    agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
    for d in per.values():
        a = agg[d['name']]
        a[0] += 1
        a[1] += d.get('ms', 0.0)
        a[2] += d.get('rd', 0.0)
        a[3] += d.get('wr', 0.0)
    tot = sum(v[1] for v in agg.values())
    n = sum(v[0] for v in agg.values())
    print(f'{n} launches, {tot:.1f} ms of kernel time (under ncu: cold-cache, serialised -- compare SHARES, not absolutes)')
    print(f'{"ms":>10} {"share":>6} {"n":>6} {"dram rd MB":>11} {"dram wr MB":>11} {"GB/s":>7}  kernel')
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        gbs = (v[2] + v[3]) / v[1] if v[1] else 0.0
        print(f'{v[1]:10.2f} {100 * v[1] / tot:5.1f}% {v[0]:6d} {v[2]:11.1f} {v[3]:11.1f} {gbs:7.0f}  {k}')


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30, int(sys.argv[3]) if len(sys.argv) > 3 else 0,
         int(sys.argv[4]) if len(sys.argv) > 4 else 1 << 60)
