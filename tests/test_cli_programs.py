"""The reference's CLI integration programs (`zokrates_cli/tests/code/*.zok` with their `*.arguments.json` and
`*.expected.witness.json`, driven by zokrates_cli/tests/integration.rs:84-398) through the native front door.

The compiler cannot run here, so each program is flattened by hand into the statement shapes the reference's flattener emits for it
(a solver directive followed by the constraints that pin its outputs: `ConditionEq` for `==`, `Bits` for comparisons, `Div` for `/`),
written as a compiled-program file and executed by `zkb_prog_compute_witness`.  The expected return values are the reference's
(copied from the `*.expected.witness.json` files named per case — small known answers, test infrastructure).  The host mirror of the
interpreter must give the same witness file byte for byte.  CPU tier: host-emulation build; -m gpu: libzkb200.so."""
import pytest

from zokrates_b200 import ir, zir
from zokrates_b200._lib import Context
from zokrates_b200.curves import curve as get_curve
from zokrates_b200.ir import Constraint, Directive, LinComb, Parameter, Prog, QuadComb, Variable, Witness

V = Variable
R = get_curve("bn128").r


class Flat:
    """Tiny flattener: fresh variables, definitions, and the directive + constraint gadgets."""

    def __init__(self, n_args, private=()):
        self.args = [V.new(i) for i in range(n_args)]
        self.private = set(private)
        self.n = n_args
        self.st = []

    def new(self):
        self.n += 1
        return V.new(self.n - 1)

    @staticmethod
    def lc(*terms):
        return LinComb([(v, k % R) for v, k in terms])

    def define(self, left, right):
        """t = left * right (left, right: LinComb)"""
        t = self.new()
        self.st.append(Constraint(QuadComb(left, right), LinComb.from_var(t)))
        return t

    def assert_eq(self, left, right, lin):
        self.st.append(Constraint(QuadComb(left, right), lin))

    def neq(self, diff: LinComb):
        """b = (diff != 0) the way `==` is flattened: ConditionEq gives (b, inv); diff * inv = b; (1 - b) * diff = 0."""
        b, inv = self.new(), self.new()
        self.st.append(Directive([QuadComb(diff, LinComb.one())], [b, inv], "ConditionEq"))
        self.st.append(Constraint(QuadComb(diff, LinComb.from_var(inv)), LinComb.from_var(b)))
        self.st.append(Constraint(QuadComb(self.lc((V.one(), 1), (b, -1)), diff), LinComb.zero()))
        return b

    def bits(self, value: LinComb, width):
        """big-endian bits of `value` (Bits directive), booleanity and recomposition constraints"""
        bs = [self.new() for _ in range(width)]
        self.st.append(Directive([QuadComb(value, LinComb.one())], bs, "Bits", width))
        for t in bs:
            self.st.append(Constraint(QuadComb(LinComb.from_var(t), LinComb.from_var(t)), LinComb.from_var(t)))
        self.st.append(Constraint(QuadComb(self.lc(*[(t, 1 << (width - 1 - i)) for i, t in enumerate(bs)]), LinComb.one()), value))
        return bs

    def ret(self, k, value: LinComb):
        self.st.append(Constraint(QuadComb(value, LinComb.one()), LinComb.from_var(V.public(k))))

    def prog(self, n_ret):
        params = [Parameter(a, i in self.private) for i, a in enumerate(self.args)]
        return Prog(params, n_ret, self.st, "bn128")


def simple_add():
    f = Flat(2); a, b = f.args
    f.ret(0, f.lc((a, 1), (b, 1)))
    return f.prog(1), [1, 2], [3]                                  # simple_add.expected.witness.json


def simple_mul():
    f = Flat(3); a, b, c = f.args
    t = f.define(LinComb.from_var(a), LinComb.from_var(b))
    u = f.define(LinComb.from_var(t), LinComb.from_var(c))
    f.ret(0, LinComb.from_var(u))
    return f.prog(1), [2, 3, 4], [24]                              # simple_mul.expected.witness.json


def arithmetics():
    f = Flat(2); a, b = f.args
    sq = f.define(f.lc((b, 1), (a, 1)), f.lc((b, 1), (a, 1)))
    f.ret(0, f.lc((a, 3), (sq, 1)))
    return f.prog(1), [1, 2], [12]                                 # arithmetics.expected.witness.json


def conditional(arg, want):
    f = Flat(1); a, = f.args
    ne = f.neq(f.lc((a, 1), (V.one(), -1)))                        # a == 1 ? 1 : 0   ->  1 - (a != 1)
    f.ret(0, f.lc((V.one(), 1), (ne, -1)))
    return f.prog(1), [arg], [want]                                # conditional_true / conditional_false .expected.witness.json


def no_return():
    f = Flat(2); a, b = f.args
    f.assert_eq(LinComb.from_var(a), LinComb.one(), LinComb.from_var(b))
    return f.prog(0), [1, 1], []                                   # no_return.expected.witness.json: {}


def return_array():
    f = Flat(8, private=range(8)); xs = f.args                     # a[3], b, c[4]
    order = [xs[3], xs[0], xs[1], xs[2], xs[4], xs[5], xs[6], xs[7]]
    for k, v in enumerate(order):
        f.ret(k, LinComb.from_var(v))
    return f.prog(8), [1, 1, 1, 2, 3, 3, 3, 3], [2, 1, 1, 1, 3, 3, 3, 3]   # return_array.expected.witness.json


def taxation():
    """x = wealth < debt ? 0 : wealth - debt.  The comparison is a bit decomposition of 2^64 + wealth - debt (inputs are small):
    its top bit says wealth >= debt."""
    f = Flat(2); debt, wealth = f.args
    bs = f.bits(f.lc((V.one(), 1 << 64), (wealth, 1), (debt, -1)), 65)
    ge = bs[0]
    x = f.define(LinComb.from_var(ge), f.lc((wealth, 1), (debt, -1)))
    f.ret(0, LinComb.from_var(x))
    return f.prog(1), [15, 12], [0]                                # taxation.expected.witness.json


def n_choose_k():
    """fac(x): 99 rounds of `f = counter == x ? f : f * i; counter = counter == x ? counter : counter + 1`; n! / (k! (n-k)!)."""
    f = Flat(2); n, k = f.args

    def fac(x: LinComb):
        acc, counter = LinComb.one(), LinComb.zero()
        for i in range(1, 100):
            ne = f.neq(LinComb(counter.value + [(v, (-c) % R) for v, c in x.value]))          # counter != x
            acc = LinComb.from_var(f.define(acc, f.lc((V.one(), 1), (ne, i - 1))))             # acc * (1 + (i - 1) ne)
            counter = LinComb(counter.value + [(ne, 1)])
        return acc
    fn, fk, fnk = fac(LinComb.from_var(n)), fac(LinComb.from_var(k)), fac(f.lc((n, 1), (k, -1)))
    den = f.define(fk, fnk)
    q = f.new()
    f.st.append(Directive([QuadComb(fn, LinComb.one()), QuadComb(LinComb.from_var(den), LinComb.one())], [q], "Div"))
    f.assert_eq(LinComb.from_var(q), LinComb.from_var(den), fn)
    f.ret(0, LinComb.from_var(q))
    return f.prog(1), [5, 1], [5]                                  # n_choose_k.expected.witness.json


def multidim_update():
    f = Flat(4); a = f.args
    for k in range(3):
        f.ret(k, LinComb.from_var(a[k]))
    f.ret(3, f.lc((V.one(), 42)))
    return f.prog(4), [0, 0, 0, 0], [0, 0, 0, 42]                  # multidim_update.expected.witness.json


CASES = [simple_add, simple_mul, arithmetics, lambda: conditional(1, 1), lambda: conditional(0, 0), no_return, return_array, taxation,
         n_choose_k, multidim_update]
NAMES = ["simple_add", "simple_mul", "arithmetics", "conditional_true", "conditional_false", "no_return", "return_array", "taxation",
         "n_choose_k", "multidim_update"]


def _check(lib, case):
    prog, inputs, want = CASES[case]()
    ref = ir.Interpreter().execute(prog, inputs)
    assert ref.return_values() == want, NAMES[case]                 # the mirror reproduces the reference's expected outputs
    ctx = Context(0, 0, lib)
    h = ctx.prog_load(zir.write_prog(prog))
    try:
        wit = ctx.prog_compute_witness(h, inputs)
        assert wit == ref.write(), NAMES[case]                      # same witness FILE as the interpreter mirror
        assert Witness.read(wit, "bn128").return_values() == want
        pub = ctx.prog_public_inputs(h)
        assert pub[len(pub) - len(want):] == want
    finally:
        ctx.prog_free(h)


@pytest.mark.parametrize("case", range(len(CASES)), ids=NAMES)
def test_cli_programs_emu(case, emu_lib):
    _check(emu_lib, case)


@pytest.mark.gpu
@pytest.mark.parametrize("case", range(len(CASES)), ids=NAMES)
def test_cli_programs_gpu(case, gpu_lib):
    _check(gpu_lib, case)
