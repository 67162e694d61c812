"""DRAM bytes per launch of the dense kernels from an ncu launch list (gpu__time_duration.sum, dram__bytes_read.sum,
dram__bytes_write.sum) -> profiles/<tag>_traffic.json (bench.py reads the newest one for roofline.traffic):
    python tools/make_traffic.py launches.csv first_id end_id tag "how the list was produced"
"""
import collections
import csv
import json
import sys

path, first, end, tag, how = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4], sys.argv[5]
per = collections.defaultdict(lambda: [0, 0.0])         # kernel -> [launches, bytes]
seen = set()
for row in csv.DictReader(l for l in open(path) if not l.startswith('==')):
    i = int(row['ID'])
    if i < first or i >= end or 'dram__bytes' not in row['Metric Name']:
        continue
    v = float(row['Metric Value'].replace(',', ''))
    u = row['Metric Unit'].lower()
    v *= 1e3 if u.startswith('k') else 1e6 if u.startswith('m') else 1e9 if u.startswith('g') else 1.0
    name = row['Kernel Name']
    key = next((k for k in ('spconv_tc_kernel', 'linear_tc_kernel', 'swformer_mlp_tc_kernel', 'mlp_chain_tc_kernel',
                            'window_attention_tc_kernel', 'qkv_proj_tc_kernel') if k in name), None)
    if key is None:
        continue
    per[key][1] += v
    if (i, key) not in seen:
        seen.add((i, key))
        per[key][0] += 1
out = {}
for k, (n, b) in per.items():
    rec = {'launches_per_step': n, 'dram_bytes_per_step': b, 'dram_bytes_per_launch': b / max(n, 1)}
    if k == 'spconv_tc_kernel':
        out.update({'kernel': k, **rec})
    else:
        out[k] = rec
out['source'] = how
json.dump(out, open(f'profiles/{tag}_traffic.json', 'w'), indent=1)
print(json.dumps(out, indent=1))
