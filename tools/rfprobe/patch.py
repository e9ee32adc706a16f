"""Rewrites the register fields / reuse flags / stall counts of the FFMA2 (or FFMA) instructions of tools/rfprobe/base.cubin.

    python patch.py            -> writes variants/*.cubin (see VARIANTS below)

Encoding (sm_70+ 128-bit SASS): Rd bits[16:24) Ra [24:32) Rb [32:40) Rc [64:72); stall [105:109) yield 109; reuse flags
bits [122:126) (bit 122 = operand slot a, 123 = b, 124 = c).
"""
import os
import re
import struct
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def text_section(cubin, kernel):
    """(file offset, size) of .text.<kernel> in the ELF64 image"""
    e_shoff = struct.unpack_from('<Q', cubin, 0x28)[0]
    e_shentsize, e_shnum, e_shstrndx = struct.unpack_from('<HHH', cubin, 0x3A)
    def sh(i):
        return struct.unpack_from('<IIQQQQIIQQ', cubin, e_shoff + i * e_shentsize)
    stroff = sh(e_shstrndx)[4]
    for i in range(e_shnum):
        name_off, _, _, _, off, size = sh(i)[:6]
        end = cubin.index(b'\0', stroff + name_off)
        if cubin[stroff + name_off:end].decode() == '.text.' + kernel:
            return off, size
    raise KeyError(kernel)


def instr_addrs(path, kernel, mnemonic):
    txt = subprocess.run(['cuobjdump', '-sass', '-fun', kernel, path], capture_output=True, text=True, check=True).stdout
    out = []
    for line in txt.splitlines():
        m = re.match(r'\s+/\*([0-9a-f]{4,6})\*/\s+(\S+)', line)
        if m and m.group(2) == mnemonic:
            out.append(int(m.group(1), 16))
    return out


def patch(base, kernel, mnemonic, ni, gen):
    """gen(j) -> dict(d=, a=, b=, c=, reuse=, stall=) for the j-th instruction of the loop body"""
    img = bytearray(base)
    off, _ = text_section(base, kernel)
    addrs = instr_addrs(os.path.join(HERE, 'base.cubin'), kernel, mnemonic)
    assert len(addrs) % ni == 0, (len(addrs), ni)
    for n, a in enumerate(addrs):
        spec = gen(n % ni)
        lo, hi = struct.unpack_from('<QQ', img, off + a)
        def setf(v, pos, width, val):
            mask = ((1 << width) - 1) << pos
            return (v & ~mask) | ((val << pos) & mask)
        lo = setf(lo, 16, 8, spec['d']); lo = setf(lo, 24, 8, spec['a']); lo = setf(lo, 32, 8, spec['b'])
        hi = setf(hi, 0, 8, spec['c'])
        hi = setf(hi, 122 - 64, 4, spec.get('reuse', 0))
        if 'stall' in spec:
            hi = setf(hi, 105 - 64, 4, spec['stall'])
        if 'yield_' in spec:
            hi = setf(hi, 109 - 64, 1, spec['yield_'])
        struct.pack_into('<QQ', img, off + a, lo, hi)
    return bytes(img)


def main():
    from variants import VARIANTS
    base = open(os.path.join(HERE, 'base.cubin'), 'rb').read()
    os.makedirs(os.path.join(HERE, 'variants'), exist_ok=True)
    for name, (kernel, mnemonic, ni, gen) in VARIANTS.items():
        out = os.path.join(HERE, 'variants', f'{kernel}__{name}.cubin')
        open(out, 'wb').write(patch(base, kernel, mnemonic, ni, gen))
    print(len(VARIANTS), 'variants written')


if __name__ == '__main__':
    main()
