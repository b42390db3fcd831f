#!/usr/bin/env python
"""Top SASS instructions by stall samples from `ncu -i X.ncu-rep --page source --csv` output (one kernel).
usage: ncu -i rep --page source --csv --kernel-name regex:NAME --launch-count 1 > src.csv; python tools/ncu_source_top.py src.csv [N]"""
import csv
import sys


def load(path):
    rows = list(csv.reader(open(path)))
    his = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
    hi = his[0]
    end = his[1] - 1 if len(his) > 1 else len(rows)
    return rows[hi], [r for r in rows[hi + 1:end] if len(r) == len(rows[hi])]


def main():
    h, data = load(sys.argv[1])
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    ia, isamp, iex = h.index('Source'), h.index('# Samples'), h.index('Instructions Executed')
    stalls = [c for c in h if c.startswith('stall_') and 'Not Issued' not in c]
    num = lambda r, i: int(r[i] or 0)
    tot = sum(num(r, isamp) for r in data)
    totex = sum(num(r, iex) for r in data)
    print('samples', tot, 'warp instructions executed', totex, 'sass lines', len(data))
    agg = {s: sum(num(r, h.index(s)) for r in data) for s in stalls}
    for s, v in sorted(agg.items(), key=lambda x: -x[1])[:10]:
        print('  %-24s %8d %5.1f%%' % (s, v, 100.0 * v / max(tot, 1)))
    order = sorted(range(len(data)), key=lambda k: -num(data[k], isamp))[:n]
    for k in order:
        r = data[k]
        why = sorted(((num(r, h.index(s)), s) for s in stalls), reverse=True)[:2]
        print('%5d %6d %8d  %-70s %s' % (k, num(r, isamp), num(r, iex), r[ia][:70], ' '.join('%s=%d' % (s[6:], v) for v, s in why if v)))


if __name__ == '__main__':
    main()
