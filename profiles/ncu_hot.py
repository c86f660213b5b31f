#!/usr/bin/env python
"""Top stall-sample SASS instructions of a kernel from an .ncu-rep (ncu -i <rep> --page source --csv)."""
import csv, subprocess, sys

def main(rep, top=25, kernel_idx=0):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    # split per kernel: a block starts with a "Kernel Name" row followed by a header row
    blocks, cur = [], None
    for r in rows:
        if r and r[0] == 'Kernel Name':
            cur = {'name': r[1], 'rows': []}
            blocks.append(cur)
        elif cur is not None:
            cur['rows'].append(r)
    b = blocks[kernel_idx]
    hdr, data = b['rows'][0], b['rows'][1:]
    si, ai, ei = hdr.index('Warp Stall Sampling (All Samples)'), hdr.index('Source'), hdr.index('Instructions Executed')
    tot = sum(int(r[si]) for r in data if len(r) > si)
    print(b['name'][:100], 'total samples', tot, 'sass lines', len(data))
    idx = {id(r): i for i, r in enumerate(data)}
    for r in sorted(data, key=lambda r: -int(r[si]))[:top]:
        print(f"{100*int(r[si])/tot:5.1f}%  #{idx[id(r)]:5d}  exec {int(r[ei]):>10d}  {r[ai].strip()[:90]}")

if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25, int(sys.argv[3]) if len(sys.argv) > 3 else 0)
