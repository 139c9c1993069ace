#!/usr/bin/env python
"""Per-kernel SASS opcode census of libndb200.so (cuobjdump -sass): the Blackwell-specific instructions each kernel contains
(UTCHMMA = tcgen05.mma, UTMALDG / UTMASTG = TMA load / store, LDTM / STTM = tcgen05.ld / st, UTCBAR = tcgen05.commit,
SYNCS = mbarrier, MUFU = special-function unit) plus registers / spills from the ptxas log of the same build.

    python scripts/sass_census.py > profiles/r02_sass_census.txt
"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'nice-diffusion_b200', 'libndb200.so')
OPS = ['UTCHMMA', 'UTMALDG', 'UTMASTG', 'UTMAPF', 'LDTM', 'STTM', 'UTCBAR', 'SYNCS', 'MUFU', 'HMMA', 'FFMA', 'LDG', 'STG', 'LDS',
       'STS', 'LDL', 'STL', 'R2UR', 'BAR', 'ACQBULK', 'UCGABAR']


def main():
    out = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True, check=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.search(r'Function : (\S+)', line)
        if m:
            name = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = kernels.setdefault(re.sub(r'\(.*', '', name), collections.Counter())
            continue
        m = re.search(r'^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)', line)
        if m and cur is not None:
            cur[m.group(1)] += 1
            cur['_total'] += 1
    regs = {}
    log = os.path.join(ROOT, 'nice-diffusion_b200', 'csrc', 'build.log')
    if os.path.exists(log):
        name = None
        for line in open(log):
            m = re.search(r"Compiling entry function '(\S+)'", line)
            if m:
                name = re.sub(r'\(.*', '', subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip())
            m = re.search(r'Used (\d+) registers', line)
            if m and name:
                regs.setdefault(name, [m.group(1), ''])[0] = m.group(1)
            m = re.search(r'(\d+) bytes spill stores', line)
            if m and name:
                regs.setdefault(name, ['', ''])[1] = m.group(1)
    print('SASS opcode census of {} (sm_100a)'.format(os.path.relpath(LIB, ROOT)))
    print('{:78s} {:>6s} {:>5s} {:>5s} '.format('kernel', 'instr', 'regs', 'spill') + ' '.join('{:>7s}'.format(o) for o in OPS))
    for name, c in kernels.items():
        r = regs.get(name, ['', ''])
        print('{:78s} {:6d} {:>5s} {:>5s} '.format(name[:78], c['_total'], r[0], r[1]) + ' '.join('{:7d}'.format(c[o]) for o in OPS))


if __name__ == '__main__':
    main()
