#!/usr/bin/env python
"""Sum per-kernel ncu --csv metrics (dram bytes, duration) grouped by kernel name."""
import csv, sys, collections
def main(path):
    rows = []
    with open(path, newline='') as f:
        lines = [l for l in f if not l.startswith('==')]
    rd = csv.DictReader(lines)
    agg = collections.defaultdict(lambda: collections.defaultdict(float))
    cnt = collections.Counter()
    seen = set()
    for r in rd:
        k = r.get('Kernel Name', '?')[:60]
        m = r.get('Metric Name'); v = r.get('Metric Value', '0').replace(',', '')
        u = r.get('Metric Unit', '')
        try: v = float(v)
        except ValueError: continue
        scale = {'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'byte': 1, 'usecond': 1e-6, 'msecond': 1e-3, 'nsecond': 1e-9, 'ns': 1e-9, 'us': 1e-6, 'ms': 1e-3, 'second': 1}.get(u, 1)
        agg[k][m] += v * scale
        if (r.get('ID'), k) not in seen:
            seen.add((r.get('ID'), k)); cnt[k] += 1
    print(path)
    for k, d in agg.items():
        print(f"  {k:60s} n={cnt[k]:4d} read={d.get('dram__bytes_read.sum',0)/1e9:8.3f} GB write={d.get('dram__bytes_write.sum',0)/1e9:8.3f} GB "
              f"time={d.get('gpu__time_duration.sum',0)*1e3:8.3f} ms")
if __name__ == '__main__':
    for p in sys.argv[1:]:
        try: main(p)
        except Exception as e: print(p, 'ERR', e)
