#!/usr/bin/env python
"""Top source lines of a kernel by warp-stall samples, from an .ncu-rep captured with --import-source on
(read here, no GPU needed):   python tools/ncu_hot_lines.py report.ncu-rep [N]"""
import csv, io, subprocess, sys, collections

def main(path, top=30):
    out = subprocess.run(['ncu', '-i', path, '--page', 'source', '--csv', '--print-source', 'cuda,sass'],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    fname, hdr = None, None
    agg = collections.defaultdict(lambda: collections.Counter())
    text = {}
    cur = None
    for r in rows:
        if not r: continue
        if r[0] in ('File Name', 'File Path'): fname = r[1].split('/')[-1]; continue
        if r[0] in ('Kernel Name', 'Function Name'): continue
        if r[0] == 'Line No': hdr = r; continue
        if hdr is None or len(r) < len(hdr) - 2: continue
        if r[0] != '':                       # a CUDA source line; its SASS rows follow with an empty first column
            cur = (fname, int(r[0])); text[cur] = r[1].strip()[:110]
        d = dict(zip(hdr[4:], r[4:]))
        if r[2] == '' and r[0] != '': continue      # the source row itself repeats the sum of its SASS rows
        try: n = int(d.get('# Samples', '0') or 0)
        except ValueError: n = 0
        if cur is None or n == 0: continue
        agg[cur]['samples'] += n
        for k, v in d.items():
            if k.startswith('stall_') and not k.endswith('(Not Issued)') and v not in ('', '0'):
                agg[cur][k] += int(v)
    tot = sum(a['samples'] for a in agg.values()) or 1
    print(f"{path}: {tot} samples")
    for key, a in sorted(agg.items(), key=lambda kv: -kv[1]['samples'])[:top]:
        st = sorted(((k, v) for k, v in a.items() if k != 'samples'), key=lambda kv: -kv[1])[:3]
        print(f"{100 * a['samples'] / tot:5.1f} %  {key[0]}:{key[1]:<4d} {text.get(key, '')}\n         " +
              ', '.join(f"{k[6:]} {100 * v / a['samples']:.0f}%" for k, v in st))

if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30)
