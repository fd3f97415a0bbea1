"""Rewrite profiles/traffic.json from the *_metrics.txt summaries of one round tag:
    python tools/restamp_traffic.py r4c <csrc sha the captures were taken on>
DRAM bytes per launch = dram__bytes_read.sum + dram__bytes_write.sum of the `ncu --set full`
capture (tools/ncu_report.sh); bench.py quotes an entry only while the sha matches the tree."""
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, sha = sys.argv[1], sys.argv[2]
UNIT = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
ENTRIES = {
    'salamander_swim:65536:16': ('fast_kernel_lean_slim_wpb7_65536', 'fb_fast_kernel<32,1,1,1> x 7 warps per block'),
    'salamander_swim:8192:16': ('fast_split_kernel_8192', 'fb_fast_split_kernel<1> x 3 warps per 32 envs'),
    'salamander:4096:16': ('fastc_split_kernel_ground4096', 'fb_fastc_split_kernel<32,1> x 3 warps per 32 envs'),
    'centipede:8192:16': ('fastc_split_kernel_centipede8192', 'fb_fastc_split_kernel<32,1> x 4 warps per 32 envs, two waves'),
}
path = os.path.join(ROOT, 'profiles', 'traffic.json')
with open(path) as f:
    tj = json.load(f)
tj['csrc_sha'] = sha
for key, (stem, kernel) in ENTRIES.items():
    source = f'profiles/{tag}_{stem}_metrics.txt'
    if not os.path.exists(os.path.join(ROOT, source)):
        print('missing', source)
        continue
    total = 0.0
    for line in open(os.path.join(ROOT, source)):
        m = re.match(r'dram__bytes_(read|write)\.sum \[(\w+)\] ([0-9.eE+-]+)', line)
        if m:
            total += float(m.group(3))*UNIT[m.group(2)]
    tj['by_workload'][key] = {'dram_bytes_per_launch': total, 'kernel': kernel, 'source': source, 'csrc_sha': sha}
    print(key, total)
with open(path, 'w') as f:
    json.dump(tj, f, indent=1)
    f.write('\n')
