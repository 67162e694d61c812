"""Per-kernel counts of the SASS mnemonics that prove tcgen05 / TMEM / TMA / mbarrier use (B200_PROFILING.md):
    python tools/sass_mnemonics.py [libos3d.so] > profiles/<round>_sass_mnemonics.txt"""
import collections
import os
import re
import subprocess
import sys

so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                        'openseg3d_b200', 'libos3d.so')
WANT = ['UTCHMMA', 'UTCBAR', 'UTCATOMSWS', 'LDTM', 'STTM', 'UTMALDG', 'UBLKCP', 'LDGSTS', 'SYNCS', 'ELECT', 'HMMA']
BLACKWELL = ['UTCHMMA', 'LDTM', 'STTM', 'UTMALDG', 'UBLKCP']
out = subprocess.run(['cuobjdump', '-sass', so], capture_output=True, text=True).stdout
kern, counts = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.match(r'\s*Function : (\S+)', line)
    if m:
        kern = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r'\(.*', '', kern)
        counts[kern] = collections.Counter()
        continue
    m = re.match(r'\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)', line)
    if m and kern:
        op = m.group(1)
        for w in WANT:
            if op == w or (w in ('LDTM', 'STTM', 'UTCBAR', 'SYNCS') and op.startswith(w)):
                counts[kern][w] += 1
print('kernels of libos3d.so that use tcgen05 (UTCHMMA), tensor memory (LDTM / STTM), TMA (UTMALDG) or bulk copies (UBLKCP);')
print('static instruction counts from cuobjdump -sass (sm_100a).  HMMA would be the legacy mma.sync path: none.')
print(f'{"kernel":58s} ' + ' '.join(f'{w:>8s}' for w in WANT))
for k, c in counts.items():
    if sum(c[w] for w in BLACKWELL):
        print(f'{k[:58]:58s} ' + ' '.join(f'{c[w]:8d}' for w in WANT))
