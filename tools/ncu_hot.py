"""Hottest SASS lines of one `ncu --page source --csv` export: instructions executed and stall samples.

Usage: ncu -i X.ncu-rep --page source --csv > src.csv; python tools/ncu_hot.py src.csv [top]
"""
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    h = rows[1]
    ia, isrc, isamp, iinst = h.index('Address'), h.index('Source'), h.index('# Samples'), h.index('Instructions Executed')
    body = [r for r in rows[2:] if len(r) > iinst]
    tot_i = sum(int(r[iinst] or 0) for r in body)
    tot_s = sum(int(r[isamp] or 0) for r in body)
    print('lines %d, warp instructions %d, samples %d' % (len(body), tot_i, tot_s))
    # contiguous regions between branch targets are easier to read than single lines: print a running index
    for n, r in enumerate(body):
        r.append(n)
    print('--- by instructions executed')
    for r in sorted(body, key=lambda r: -int(r[iinst] or 0))[:top]:
        print('%5d %10d %5.1f%% samples %6d  %s' % (r[-1], int(r[iinst]), 100.0 * int(r[iinst]) / tot_i, int(r[isamp] or 0), r[isrc].strip()[:90]))
    print('--- by stall samples')
    for r in sorted(body, key=lambda r: -int(r[isamp] or 0))[:top]:
        print('%5d %10d samples %6d %5.1f%%  %s' % (r[-1], int(r[iinst]), int(r[isamp] or 0), 100.0 * int(r[isamp] or 0) / max(tot_s, 1), r[isrc].strip()[:90]))


if __name__ == '__main__':
    main()
