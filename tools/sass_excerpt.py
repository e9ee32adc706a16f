"""SASS evidence of the shipped 2D cluster kernels (no GPU needed): python tools/sass_excerpt.py > profiles/r02_sass_excerpt.txt

Per kernel instantiation: registers / spills (ptxas.log), counts of the Blackwell / Hopper+ mnemonics that carry the design
(UTMALDG = TMA tensor load, UTMAPF = TMA L2 prefetch, STAS = st.async to distributed shared memory, UCGABAR = cluster barrier,
SYNCS = mbarrier, LDGSTS = cp.async), the steady-state two-step loop body (instruction mix) and what the register-file model
of profiles/r02_rfprobe.txt (issue cycles = max(1, #distinct even source registers outside the reuse cache, #odd)) charges it.
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'tools'))
from sass_loop import kernel_sass, opcode   # noqa: E402

LIB = os.path.join(ROOT, 'cspn_b200', '_build', 'libcspn_b200.so')
KERNELS = [('forward, chained strips <PR=5,PC=4,NW=8,abs=0,general=0,kForward,CHAIN=1>', 'cspn2d_cluster_kernelILi5ELi4ELi8ELb0ELb0ELi0ELb1'),
           ('forward <5,4,8,0,0,kForward,CHAIN=0>', 'cspn2d_cluster_kernelILi5ELi4ELi8ELb0ELb0ELi0ELb0'),
           ('forward storing every step (backward) <5,4,8,0,1,kStoreSteps>', 'cspn2d_cluster_kernelILi5ELi4ELi8ELb0ELb1ELi1ELb0'),
           ('adjoint <5,4,8,0,1,kAdjoint>', 'cspn2d_cluster_kernelILi5ELi4ELi8ELb0ELb1ELi2ELb0')]
MNEMONICS = ['UTMALDG', 'UTMAPF', 'STAS', 'UCGABAR', 'SYNCS', 'LDGSTS', 'FFMA', 'SHFL', 'LDS', 'STS', 'MUFU', 'STG', 'LDG', 'STL', 'LDL']


def ptxas_info(pattern):
    log = open(os.path.join(ROOT, 'cspn_b200', '_build', 'ptxas.log')).read().splitlines()
    for i, l in enumerate(log):
        if 'Compiling entry function' in l and pattern in l:
            return ' '.join(x.strip() for x in log[i + 2:i + 4])
    return '?'


def rf_cycles(body):
    cached = [None] * 4
    tot = nf = fc = 0
    for _, s in body:
        s2 = re.sub(r'^@!?U?P\d+\s+', '', s)
        parts = s2.split(None, 1)
        op = parts[0]
        ops = [o.strip() for o in parts[1].split(',')] if len(parts) > 1 else []
        srcs = ops if op.startswith(('STS', 'STG', 'STAS')) else ops[1:]
        even, odd, new = set(), set(), [None] * 4
        for slot, o in enumerate(srcs[:4]):
            m = re.search(r'(?<![U\w])R(\d+)(\.reuse)?', o)
            if not m:
                continue
            r = int(m.group(1))
            w = 4 if (op.endswith('.128') and slot == len(srcs) - 1 and op.startswith(('STS', 'STAS'))) else (2 if '.64' in o else 1)
            if cached[slot] != r:
                for x in range(r, r + w):
                    (even if x % 2 == 0 else odd).add(x)
            if m.group(2):
                new[slot] = r
        cached = new
        c = max(1, len(even), len(odd))
        tot += c
        if op.startswith('FFMA'):
            nf += 1
            fc += c
    return tot, nf, fc


def main():
    print(__doc__.strip().splitlines()[0].split(':')[0] + f' -- {os.path.relpath(LIB, ROOT)}\n')
    for title, pat in KERNELS:
        ins = kernel_sass(LIB, pat)
        if not ins:
            print(f'== {title}: not in the library\n')
            continue
        print(f'== {title}\n   {len(ins)} instructions; {ptxas_info(pat)}')
        cnt = collections.Counter()
        for _, s in ins:
            op = opcode(s)
            for m in MNEMONICS:
                if op.startswith(m):
                    cnt[m] += 1
        print('   ' + '  '.join(f'{m} {cnt[m]}' for m in MNEMONICS))
        first = {}
        for a, s in ins:
            for m in ('UTMALDG', 'UTMAPF', 'STAS', 'UCGABAR', 'LDGSTS'):
                if opcode(s).startswith(m) and m not in first:
                    first[m] = f'/*{a:05x}*/ {s}'
        for m, l in first.items():
            print(f'   first {m:8s} {l}')
        # steady-state loop: the shortest loop holding >= 300 FFMAs
        addr = {a: i for i, (a, _) in enumerate(ins)}
        loops = []
        for i, (a, s) in enumerate(ins):
            m = re.search(r'BRA(?:\.U)?(?:\.ANY)?\s+(?:!?U?P\d+,\s*)?(0x[0-9a-f]+)', s)
            if m and int(m.group(1), 16) in addr and addr[int(m.group(1), 16)] < i:
                loops.append((addr[int(m.group(1), 16)], i))
        cand = sorted((l for l in loops if sum('FFMA' in ins[k][1] for k in range(l[0], l[1] + 1)) >= 300), key=lambda l: l[1] - l[0])
        if cand:
            body = ins[cand[0][0]:cand[0][1] + 1]
            mix = collections.Counter(opcode(s) for _, s in body)
            tot, nf, fc = rf_cycles(body)
            print(f'   two-step loop body: {len(body)} instructions: ' + ', '.join(f'{k} {v}' for k, v in mix.most_common(12)))
            print(f'   register-file model: {tot} issue cycles per warp and two steps ({nf} FFMA cost {fc}, the other {len(body) - nf} '
                  f'instructions {tot - fc}); two warps per sub-partition -> {tot} cycles per step')
        print()


if __name__ == '__main__':
    main()
