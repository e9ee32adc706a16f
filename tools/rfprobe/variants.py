"""Operand patterns of the register-file probe.  Registers R16..R159 are free inside the loop (R0, R8, R9 hold the thread
id and the start clock).  An FFMA2 operand is the pair (R2n, R2n+1), named by n; 'cls' picks registers by n % m == r."""
NI = 96
PAIRS = list(range(8, 77))          # pair indices n: R16..R153 (the two highest allocated registers are not addressable)
REGS = list(range(16, 154))


def pick(pool, j, stride=1):
    return pool[(j * stride) % len(pool)]


def pairs_where(m, r):
    return [n for n in PAIRS if n % m == r]


def regs_where(m, r):
    return [x for x in REGS if x % m == r]


def f2(a_pool, b_pool, c_pool, b_group=1, reuse_b=False, stall=None, c_count=16, a_eq_b=False):
    """FFMA2 d=c, a*b+c; b changes every b_group instructions; accumulators cycle over c_count pairs of c_pool"""
    def gen(j):
        c = 2 * c_pool[j % min(c_count, len(c_pool))]
        a = 2 * pick(a_pool, j, 7)
        b = 2 * pick(b_pool, j // b_group, 5)
        if a_eq_b:
            b = a
        # make sure the three are distinct pairs unless asked otherwise
        s = dict(d=c, a=a, b=b, c=c)
        if reuse_b and (j % b_group) != b_group - 1:
            s['reuse'] = 0b0010
        if stall is not None:
            s['stall'] = stall
        return s
    return ('k_ffma2', 'FFMA2', NI, gen)


def f1(a_pool, b_pool, c_pool, b_group=1, reuse_b=False, stall=None, c_count=32, a_eq_b=False):
    def gen(j):
        c = c_pool[j % min(c_count, len(c_pool))]
        a = pick(a_pool, j, 7)
        b = pick(b_pool, j // b_group, 5)
        if a_eq_b:
            b = a
        s = dict(d=c, a=a, b=b, c=c)
        if reuse_b and (j % b_group) != b_group - 1:
            s['reuse'] = 0b0010
        if stall is not None:
            s['stall'] = stall
        return s
    return ('k_ffma', 'FFMA', NI, gen)


P = pairs_where
R = regs_where
# disjoint thirds so that a, b, c never coincide by accident
def third(pool, i):
    k = len(pool) // 3
    return pool[i * k:(i + 1) * k]


VARIANTS = {
    # ---- FFMA2: parity of the pair index n (register number / 2) ----
    'p2_a0_b0_c0': f2(third(P(2, 0), 0), third(P(2, 0), 1), third(P(2, 0), 2)),
    'p2_a0_b0_c1': f2(third(P(2, 0), 0), third(P(2, 0), 1), third(P(2, 1), 2)),
    'p2_a0_b1_c1': f2(third(P(2, 0), 0), third(P(2, 1), 1), third(P(2, 1), 2)),
    'p2_a1_b0_c0': f2(third(P(2, 1), 0), third(P(2, 0), 1), third(P(2, 0), 2)),
    # n % 4 classes
    'p4_a0_b1_c2': f2(third(P(4, 0), 0), third(P(4, 1), 1), third(P(4, 2), 2), c_count=6),
    'p4_a0_b0_c0': f2(third(P(4, 0), 0), third(P(4, 0), 1), third(P(4, 0), 2), c_count=6),
    'p4_a0_b2_c0': f2(third(P(4, 0), 0), third(P(4, 2), 1), third(P(4, 0), 2), c_count=6),
    # operand reuse: b fixed for 8 instructions, with and without the .reuse flag
    'p2_a0_b0_c0_bgroup8_reuse': f2(third(P(2, 0), 0), third(P(2, 0), 1), third(P(2, 0), 2), b_group=8, reuse_b=True),
    'p2_a0_b0_c0_bgroup8_noflag': f2(third(P(2, 0), 0), third(P(2, 0), 1), third(P(2, 0), 2), b_group=8, reuse_b=False),
    'p2_a0_b0_c1_bgroup8_reuse': f2(third(P(2, 0), 0), third(P(2, 0), 1), third(P(2, 1), 2), b_group=8, reuse_b=True),
    'p2_a0_b1_c1_bgroup8_reuse': f2(third(P(2, 0), 0), third(P(2, 1), 1), third(P(2, 1), 2), b_group=8, reuse_b=True),
    'p2_a0_b1_c0_bgroup8_reuse': f2(third(P(2, 0), 0), third(P(2, 1), 1), third(P(2, 0), 2), b_group=8, reuse_b=True),
    # two distinct pairs only (a == b)
    'p2_a0_eqb_c0': f2(third(P(2, 0), 0), third(P(2, 0), 1), third(P(2, 0), 2), a_eq_b=True),
    'p2_a0_eqb_c1': f2(third(P(2, 0), 0), third(P(2, 0), 1), third(P(2, 1), 2), a_eq_b=True),
    # stall field forced
    'p2_a0_b0_c1_stall1': f2(third(P(2, 0), 0), third(P(2, 0), 1), third(P(2, 1), 2), stall=1),
    'p2_a0_b0_c1_stall2': f2(third(P(2, 0), 0), third(P(2, 0), 1), third(P(2, 1), 2), stall=2),
    'p2_a0_b0_c1_stall4': f2(third(P(2, 0), 0), third(P(2, 0), 1), third(P(2, 1), 2), stall=4),

    # ---- FFMA: parity of the register number ----
    's2_a0_b0_c0': f1(third(R(2, 0), 0), third(R(2, 0), 1), third(R(2, 0), 2)),
    's2_a0_b0_c1': f1(third(R(2, 0), 0), third(R(2, 0), 1), third(R(2, 1), 2)),
    's2_a0_b1_c1': f1(third(R(2, 0), 0), third(R(2, 1), 1), third(R(2, 1), 2)),
    's2_a0_b1_c0': f1(third(R(2, 0), 0), third(R(2, 1), 1), third(R(2, 0), 2)),
    # register number % 4 and % 8
    's4_a0_b1_c2': f1(third(R(4, 0), 0), third(R(4, 1), 1), third(R(4, 2), 2), c_count=12),
    's4_a0_b2_c0': f1(third(R(4, 0), 0), third(R(4, 2), 1), third(R(4, 0), 2), c_count=12),
    's4_a0_b2_c2': f1(third(R(4, 0), 0), third(R(4, 2), 1), third(R(4, 2), 2), c_count=12),
    's4_a0_b0_c0': f1(third(R(4, 0), 0), third(R(4, 0), 1), third(R(4, 0), 2), c_count=12),
    's8_a0_b4_c0': f1(third(R(8, 0), 0), third(R(8, 4), 1), third(R(8, 0), 2), c_count=6),
    's8_a0_b2_c4': f1(third(R(8, 0), 0), third(R(8, 2), 1), third(R(8, 4), 2), c_count=6),
    # reuse
    's2_a0_b0_c0_bgroup8_reuse': f1(third(R(2, 0), 0), third(R(2, 0), 1), third(R(2, 0), 2), b_group=8, reuse_b=True),
    's2_a0_b0_c1_bgroup8_reuse': f1(third(R(2, 0), 0), third(R(2, 0), 1), third(R(2, 1), 2), b_group=8, reuse_b=True),
    's2_a0_b0_c1_bgroup8_noflag': f1(third(R(2, 0), 0), third(R(2, 0), 1), third(R(2, 1), 2), b_group=8, reuse_b=False),
    's2_a0_b1_c1_bgroup8_reuse': f1(third(R(2, 0), 0), third(R(2, 1), 1), third(R(2, 1), 2), b_group=8, reuse_b=True),
    's2_a0_b0_c1_bgroup3_reuse': f1(third(R(2, 0), 0), third(R(2, 0), 1), third(R(2, 1), 2), b_group=3, reuse_b=True),
    # two distinct registers
    's2_a0_eqb_c0': f1(third(R(2, 0), 0), third(R(2, 0), 1), third(R(2, 0), 2), a_eq_b=True),
    's2_a0_eqb_c1': f1(third(R(2, 0), 0), third(R(2, 0), 1), third(R(2, 1), 2), a_eq_b=True),
    's2_a0_b0_c1_stall1': f1(third(R(2, 0), 0), third(R(2, 0), 1), third(R(2, 1), 2), stall=1),
    's2_a0_b0_c1_stall2': f1(third(R(2, 0), 0), third(R(2, 0), 1), third(R(2, 1), 2), stall=2),
}


# ---- second batch: how long does a reuse-cache entry live, which slots have one -------------------------------------------
def f1_custom(pattern):
    """pattern(j) -> (a, b, c, reuse_mask) with explicit registers"""
    def gen(j):
        a, b, c, ru = pattern(j)
        return dict(d=c, a=a, b=b, c=c, reuse=ru)
    return ('k_ffma', 'FFMA', NI, gen)


_A = third(R(2, 0), 0)      # even weights
_C = third(R(2, 1), 2)      # odd accumulators
_X = third(R(2, 0), 1)      # even x values


def _alt(flag_x, flag_y):
    def pat(j):
        x = _X[0] if j % 2 == 0 else _X[1]
        ru = (0b0010 if (j % 2 == 0 and flag_x) or (j % 2 == 1 and flag_y) else 0)
        return pick(_A, j, 7), x, _C[j % 32 % len(_C)], ru
    return pat


def _gap(gap):
    """X.reuse, then `gap` instructions with other b registers (no flag), then X again"""
    def pat(j):
        k = j % (gap + 1)
        x = _X[0] if k == 0 else _X[1 + (j % 7)]
        return pick(_A, j, 7), x, _C[j % len(_C)], (0b0010 if k == 0 else 0)
    return pat


def _slot_a_group(g):
    def pat(j):
        x = _X[(j // g) % len(_X)]
        return x, pick(_A, j, 7), _C[j % len(_C)], (0b0001 if j % g != g - 1 else 0)
    return pat


def _cross_slot(j):
    # even j: X in slot b with flag; odd j: X in slot a
    x = _X[(j // 2) % len(_X)]
    if j % 2 == 0:
        return pick(_A, j, 7), x, _C[j % len(_C)], 0b0010
    return x, pick(_A, j, 7), _C[j % len(_C)], 0


def _two_slots(j):
    # a and b both fixed for groups of 4, both flagged; c odd -> 1 read
    g = j // 4
    return _A[g % len(_A)], _X[g % len(_X)], _C[j % len(_C)], (0b0011 if j % 4 != 3 else 0)


def _flag_every(g):
    """b fixed for groups of g, EVERY instruction carries the flag (also the last of a group)"""
    def pat(j):
        return pick(_A, j, 7), _X[(j // g) % len(_X)], _C[j % len(_C)], 0b0010
    return pat


VARIANTS.update({
    'r_alt_flagX': f1_custom(_alt(True, False)),
    'r_alt_flagXY': f1_custom(_alt(True, True)),
    'r_alt_noflag': f1_custom(_alt(False, False)),
    'r_gap1': f1_custom(_gap(1)),
    'r_gap2': f1_custom(_gap(2)),
    'r_gap3': f1_custom(_gap(3)),
    'r_slot_a_group8': f1_custom(_slot_a_group(8)),
    'r_slot_a_group4': f1_custom(_slot_a_group(4)),
    'r_cross_slot': f1_custom(_cross_slot),
    'r_two_slots_group4': f1_custom(_two_slots),
    'r_flag_every_group4': f1_custom(_flag_every(4)),
    'r_flag_every_group2': f1_custom(_flag_every(2)),
    'p2_bgroup2_reuse': f2(third(P(2, 0), 0), third(P(2, 0), 1), third(P(2, 1), 2), b_group=2, reuse_b=True),
    'p2_bgroup4_reuse': f2(third(P(2, 0), 0), third(P(2, 0), 1), third(P(2, 1), 2), b_group=4, reuse_b=True),
    'p2_bgroup96_reuse': f2(third(P(2, 0), 0), third(P(2, 0), 1), third(P(2, 1), 2), b_group=96, reuse_b=True),
})
