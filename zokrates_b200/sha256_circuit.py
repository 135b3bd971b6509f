"""An R1CS for the `sha256packed` preimage program (BASELINE.json config 2), built in-house.

The reference compiles `zokrates_stdlib/stdlib/hashes/sha256/512bitPacked.zok` (four 128-bit field inputs ->
two 128-bit field outputs, SHA-256 of the 64-byte message with standard padding) with its own compiler, which
cannot run here (no Rust toolchain).  This module builds an equivalent constraint system directly
— same function, same input/output packing, the usual bit gadgets — together with its witness, so that a
realistic "almost all bits" circuit of that size class can be proven.  It is pinned FUNCTIONALLY by the
reference's KAT (`zokrates_stdlib/tests/tests/hashes/sha256/512bitPacked.json:5-16`: inputs 0,0,0,5) but is
not the byte-identical constraint system ZoKrates emits, so proving keys are not interchangeable.

Gadgets (variables are bits unless noted):
  booleanity      b * b = b
  xor             (2a) * b = a + b - c
  ch(e,f,g)       e * (f - g) = ch - g
  maj(a,b,c)      t = a * b ;  (a + b - 2t) * c = maj - t
  modular add     sum_k word_k = sum_i r_i 2^i   (one linear row, r_i boolean, top carries dropped)
"""
from __future__ import annotations

import hashlib

import numpy as np

from ._lib import fr_array
from .curves import curve as _curve
from .r1cs import R1CS

K = [0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98, 0x12835b01,
     0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc,
     0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147,
     0x06ca6351, 0x14292967, 0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85,
     0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070, 0x19a4c116, 0x1e376c08,
     0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208,
     0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2]
IV = [0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19]


class _Builder:
    """Rows as term lists; variable 0 is the constant one.  A 'word' is a list of 32 bit operands, LSB first;
    an operand is a variable index (int >= 0) or a constant bit ('c', 0/1)."""

    def __init__(self, r):
        self.r = r
        self.a, self.b, self.c = [], [], []
        self.z = [1]
        self.n_inst = 1

    def new(self, val):
        self.z.append(val % self.r)
        return len(self.z) - 1

    def row(self, a, b, c):
        self.a.append(a); self.b.append(b); self.c.append(c)

    def val(self, op):
        return op[1] if isinstance(op, tuple) else self.z[op]

    def lin(self, op, coeff=1):
        """term list for coeff * operand"""
        if isinstance(op, tuple):
            return [(0, coeff * op[1] % self.r)] if op[1] else []
        return [(op, coeff % self.r)]

    def boolean(self, v):
        self.row([(v, 1)], [(v, 1)], [(v, 1)])

    def xor(self, x, y):
        vx, vy = self.val(x), self.val(y)
        if isinstance(x, tuple) and isinstance(y, tuple):
            return ("c", vx ^ vy)
        if isinstance(x, tuple):
            x, y = y, x
        if isinstance(y, tuple):          # x xor const
            if y[1] == 0:
                return x
            out = self.new(1 - vx)         # not x: (x) * 1 = 1 - out
            self.row([(x, 1)], [(0, 1)], [(0, 1), (out, self.r - 1)])
            return out
        out = self.new(vx ^ vy)
        self.row([(x, 2)], [(y, 1)], [(x, 1), (y, 1), (out, self.r - 1)])
        return out

    def xor3(self, x, y, w):
        return self.xor(self.xor(x, y), w)

    def ch(self, e, f, g):
        ve, vf, vg = self.val(e), self.val(f), self.val(g)
        out = self.new(vg ^ (ve & (vf ^ vg)))
        self.row(self.lin(e), self.lin(f) + self.lin(g, -1), [(out, 1)] + self.lin(g, -1))
        return out

    def maj(self, x, y, w):
        vx, vy, vw = self.val(x), self.val(y), self.val(w)
        t = self.new(vx & vy)
        self.row(self.lin(x), self.lin(y), [(t, 1)])
        out = self.new((vx & vy) | (vx & vw) | (vy & vw))
        self.row(self.lin(x) + self.lin(y) + [(t, self.r - 2)], self.lin(w), [(out, 1), (t, self.r - 1)])
        return out

    def add_words(self, words, const=0):
        """sum of 32-bit words (+ constant) mod 2^32 -> fresh 32-bit word; carries are boolean-constrained."""
        total = const + sum(sum(self.val(bit) << i for i, bit in enumerate(w)) for w in words)
        nbits = max(33, (len(words) * ((1 << 32) - 1) + const).bit_length())
        res = [self.new((total >> i) & 1) for i in range(nbits)]
        for v in res:
            self.boolean(v)
        lhs = [(0, const % self.r)] if const else []
        for w in words:
            for i, bit in enumerate(w):
                lhs += self.lin(bit, 1 << i)
        self.row(lhs, [(0, 1)], [(v, (1 << i) % self.r) for i, v in enumerate(res)])
        return res[:32]


def _rotr(w, n):
    return w[n:] + w[:n]


def _shr(w, n):
    return w[n:] + [("c", 0)] * n


def _compress(bd: _Builder, state, block):
    w = list(block)
    for t in range(16, 64):
        s0 = [bd.xor3(a, b, c) for a, b, c in zip(_rotr(w[t - 15], 7), _rotr(w[t - 15], 18), _shr(w[t - 15], 3))]
        s1 = [bd.xor3(a, b, c) for a, b, c in zip(_rotr(w[t - 2], 17), _rotr(w[t - 2], 19), _shr(w[t - 2], 10))]
        w.append(bd.add_words([w[t - 16], s0, w[t - 7], s1]))
    a, b, c, d, e, f, g, h = state
    for t in range(64):
        S1 = [bd.xor3(x, y, v) for x, y, v in zip(_rotr(e, 6), _rotr(e, 11), _rotr(e, 25))]
        chv = [bd.ch(x, y, v) for x, y, v in zip(e, f, g)]
        S0 = [bd.xor3(x, y, v) for x, y, v in zip(_rotr(a, 2), _rotr(a, 13), _rotr(a, 22))]
        mj = [bd.maj(x, y, v) for x, y, v in zip(a, b, c)]
        new_e = bd.add_words([d, h, S1, chv, w[t]], K[t])
        new_a = bd.add_words([h, S1, chv, w[t], S0, mj], K[t])
        a, b, c, d, e, f, g, h = new_a, a, b, c, new_e, e, f, g
    return [bd.add_words([x, y]) for x, y in zip(state, [a, b, c, d, e, f, g, h])]


def make(curve, inputs=(0, 0, 0, 5)):
    """R1CS + assignment for `def main(private field a, b, c, d) -> field[2]` = sha256packed.  The two 128-bit
    outputs are public (instance variables 1, 2), the four inputs private.  Returns (r1cs, z, outputs)."""
    cv = _curve(curve)
    bd = _Builder(cv.r)
    digest = hashlib.sha256(b"".join(int(v).to_bytes(16, "big") for v in inputs)).digest()
    outs = (int.from_bytes(digest[:16], "big"), int.from_bytes(digest[16:], "big"))
    out_vars = [bd.new(outs[0]), bd.new(outs[1])]
    bd.n_inst = 3
    in_vars = [bd.new(v) for v in inputs]
    # unpack each input into 128 bits (big-endian, as the stdlib's unpack128)
    msg_bits = []
    for var, v in zip(in_vars, inputs):
        bits = [bd.new((v >> i) & 1) for i in range(128)]          # LSB first
        for bv in bits:
            bd.boolean(bv)
        bd.row([(var, 1)], [(0, 1)], [(bv, (1 << i) % cv.r) for i, bv in enumerate(bits)])
        msg_bits += bits[::-1]                                       # message order: MSB first
    words = [msg_bits[32 * k:32 * k + 32][::-1] for k in range(16)]  # each word LSB first
    state = [[("c", (iv >> i) & 1) for i in range(32)] for iv in IV]
    state = _compress(bd, state, words)
    pad = [0x80000000] + [0] * 14 + [512]
    state = _compress(bd, state, [[("c", (p >> i) & 1) for i in range(32)] for p in pad])
    # pack the digest into two 128-bit field elements
    dbits = []
    for wd in state:
        dbits += wd[::-1]                                            # MSB first
    for k in range(2):
        part = dbits[128 * k:128 * k + 128]
        lhs = []
        for i, bit in enumerate(part):
            lhs += bd.lin(bit, 1 << (127 - i))
        bd.row(lhs, [(0, 1)], [(out_vars[k], 1)])

    def csr(rows):
        rowptr = np.zeros(len(rows) + 1, dtype=np.uint64)
        cols, vals = [], []
        for i, row in enumerate(rows):
            for cidx, coeff in row:
                cols.append(cidx); vals.append(coeff % cv.r)
            rowptr[i + 1] = len(cols)
        return rowptr, np.array(cols, dtype=np.uint32), fr_array(vals)

    r1cs = R1CS(cv.name, len(bd.a), bd.n_inst, len(bd.z) - bd.n_inst, csr(bd.a), csr(bd.b), csr(bd.c))
    return r1cs, fr_array(bd.z), outs


# ---- the same function as a PROGRAM with solver directives (what `zokrates compile` hands to compute-witness) ----------------
class _ProgBuilder(_Builder):
    """Emits IR statements instead of R1CS rows with values: every gadget is the directive that computes its output followed by
    the constraints that pin it — `Bits` for the unpacking and the modular additions, `Xor`, `ShaCh`, `ShaAndXorAndXorAnd` for the
    bit gadgets (zokrates_interpreter/src/lib.rs:249-307) — so witness generation has to run the solvers, as for a compiled program."""

    def __init__(self, r, n_out, n_in):
        super().__init__(r)
        from .ir import Variable
        self.V = Variable
        self.st = []
        self.n_out, self.n_in = n_out, n_in

    def var(self, i):
        if i == 0:
            return self.V.one()
        if i <= self.n_out:
            return self.V.public(i - 1)
        return self.V.new(i - 1 - self.n_out)           # arguments first: _0 .. _(n_in - 1)

    def lc(self, terms):
        from .ir import LinComb
        return LinComb([(self.var(i), k % self.r) for i, k in terms])

    def new(self, val=0):
        self.z.append(0)
        return len(self.z) - 1

    def val(self, op):
        return op[1] if isinstance(op, tuple) else 0

    def row(self, a, b, c):
        from .ir import Constraint, QuadComb
        self.st.append(Constraint(QuadComb(self.lc(a), self.lc(b)), self.lc(c)))

    def directive(self, solver, inputs, outputs, arg=None):
        from .ir import Directive, QuadComb, LinComb
        self.st.append(Directive([QuadComb(self.lc(t), LinComb.one()) for t in inputs], [self.var(o) for o in outputs], solver, arg))

    def xor(self, x, y):
        if isinstance(x, tuple) and isinstance(y, tuple):
            return ("c", x[1] ^ y[1])
        if isinstance(x, tuple):
            x, y = y, x
        if isinstance(y, tuple):
            if y[1] == 0:
                return x
            out = self.new()
            self.row([(0, 1), (x, self.r - 1)], [(0, 1)], [(out, 1)])            # not x: a definition
            return out
        out = self.new()
        self.directive("Xor", [[(x, 1)], [(y, 1)]], [out])
        self.row([(x, 2)], [(y, 1)], [(x, 1), (y, 1), (out, self.r - 1)])
        return out

    def ch(self, e, f, g):
        out = self.new()
        self.directive("ShaCh", [self.lin(e), self.lin(f), self.lin(g)], [out])
        self.row(self.lin(e), self.lin(f) + self.lin(g, -1), [(out, 1)] + self.lin(g, -1))
        return out

    def maj(self, x, y, w):
        t = self.new()
        self.row(self.lin(x), self.lin(y), [(t, 1)])                               # t = x y: a definition
        out = self.new()
        self.directive("ShaAndXorAndXorAnd", [self.lin(w), self.lin(x), self.lin(y)], [out])   # x y - (2 x y - x - y) w
        self.row(self.lin(x) + self.lin(y) + [(t, self.r - 2)], self.lin(w), [(out, 1), (t, self.r - 1)])
        return out

    def add_words(self, words, const=0):
        nbits = max(33, (len(words) * ((1 << 32) - 1) + const).bit_length())
        res = [self.new() for _ in range(nbits)]
        lhs = [(0, const % self.r)] if const else []
        for w in words:
            for i, bit in enumerate(w):
                lhs += self.lin(bit, 1 << i)
        self.directive("Bits", [lhs], res[::-1], nbits)                            # big-endian outputs
        for v in res:
            self.boolean(v)
        self.row(lhs, [(0, 1)], [(v, (1 << i) % self.r) for i, v in enumerate(res)])
        return res[:32]


def make_prog(curve):
    """`def main(private field a, b, c, d) -> field[2]` = sha256packed as an IR program with solver directives."""
    from .ir import Parameter, Prog
    cv = _curve(curve)
    bd = _ProgBuilder(cv.r, 2, 4)
    out_vars = [bd.new(), bd.new()]
    in_vars = [bd.new() for _ in range(4)]
    msg_bits = []
    for var in in_vars:
        bits = [bd.new() for _ in range(128)]                                      # LSB first
        bd.directive("Bits", [[(var, 1)]], bits[::-1], 128)
        for bv in bits:
            bd.boolean(bv)
        bd.row([(var, 1)], [(0, 1)], [(bv, (1 << i) % cv.r) for i, bv in enumerate(bits)])
        msg_bits += bits[::-1]
    words = [msg_bits[32 * k:32 * k + 32][::-1] for k in range(16)]
    state = [[("c", (iv >> i) & 1) for i in range(32)] for iv in IV]
    state = _compress(bd, state, words)
    pad = [0x80000000] + [0] * 14 + [512]
    state = _compress(bd, state, [[("c", (p >> i) & 1) for i in range(32)] for p in pad])
    dbits = []
    for wd in state:
        dbits += wd[::-1]
    for k in range(2):
        lhs = []
        for i, bit in enumerate(dbits[128 * k:128 * k + 128]):
            lhs += bd.lin(bit, 1 << (127 - i))
        bd.row(lhs, [(0, 1)], [(out_vars[k], 1)])
    return Prog([Parameter.private_(bd.var(v)) for v in in_vars], 2, bd.st, cv.name)
