"""Count the SASS mnemonics that prove the Blackwell paths in the built library.

Usage: python tools/sass_counts.py [lib.so] > profiles/r02_sass_counts.txt
"""
import collections
import re
import subprocess
import sys

PATTERNS = {
    'UTCHMMA': r'\bUTCHMMA', 'UTMALDG': r'\bUTMALDG', 'UTMASTG': r'\bUTMASTG', 'LDTM': r'\bLDTM',
    'UTCBAR': r'\bUTCBAR', 'LDGSTS': r'\bLDGSTS', 'SYNCS': r'\bSYNCS', 'HMMA(legacy)': r'(?<![A-Z])HMMA',
    'UCGABAR': r'\bUCGABAR', 'ACQBULK': r'\bACQBULK', 'STG.256': r'\bSTG\.E\.(ENL2\.)?256', 'LDG.256': r'\bLDG\.E\.(ENL2\.)?256',
    'ELECT': r'\bELECT',
}


def main():
    lib = sys.argv[1] if len(sys.argv) > 1 else 'geeco_b200/libgeeco_b200.so'
    sass = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True, check=True).stdout
    pats = {k: re.compile(v) for k, v in PATTERNS.items()}
    counts = collections.defaultdict(collections.Counter)
    cur = None
    for line in sass.splitlines():
        m = re.search(r'Function : (\S+)', line)
        if m:
            cur = m.group(1)
            continue
        if cur:
            for k, p in pats.items():
                if p.search(line):
                    counts[cur][k] += 1
    names = subprocess.run(['c++filt'], input='\n'.join(counts), capture_output=True, text=True).stdout.splitlines()
    total = collections.Counter()
    rows = []
    for name, c in zip(names, counts.values()):
        total.update(c)
        rows.append((name.replace('(anonymous namespace)::', '')[:100], sorted(c.items())))
    print('# cuobjdump -sass %s (sm_100a): lines per kernel holding the mnemonics B200_PROFILING.md names.' % lib)
    print('# UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG = TMA tensor load, UTCBAR = tcgen05.commit, LDGSTS = cp.async,')
    print('# SYNCS = mbarrier ops, UCGABAR = cluster barrier, ACQBULK = griddepcontrol.wait; HMMA(legacy) = mma.sync and must be 0.')
    print('TOTAL', dict(sorted(total.items())))
    print()
    for name, c in sorted(rows):
        print('%-102s %s' % (name, dict(c)))


if __name__ == '__main__':
    main()
