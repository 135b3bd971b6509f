"""The native front door (zkb_prog_load / zkb_prog_compute_witness / zkb_prog_set_witness): C++ reader of the compiled
program file, ark-order synthesis, level schedule and the solver kernels, against the Python mirror of the reference
(zokrates_b200/zir.py, r1cs.py, ir.Interpreter — themselves pinned to the reference's KATs in tests/test_oracle_pins.py and
tests/test_zir_format.py).  CPU tier: the kernel bodies run in the host-emulation build; -m gpu: the same through libzkb200.so."""
import random

import numpy as np
import pytest

from zokrates_b200 import ir, zir
from zokrates_b200._lib import Context, ZkbError
from zokrates_b200.curves import curve as get_curve
from zokrates_b200.ir import Constraint, Directive, LinComb, Parameter, Prog, QuadComb, Variable, Witness
from zokrates_b200.r1cs import synthesize

V = Variable


def lc(*terms):
    return LinComb([(v, k) for v, k in terms])


def q(left, right=None):
    left = left if isinstance(left, LinComb) else LinComb.from_var(left)
    right = LinComb.one() if right is None else (right if isinstance(right, LinComb) else LinComb.from_var(right))
    return QuadComb(left, right)


def solver_program(curve="bn128", width=8):
    """Every simple solver once, wired the way the compiler wires them (directive, then the constraints that pin its outputs)."""
    r = get_curve(curve).r
    x, y, p = V.new(0), V.new(1), V.new(2)
    nxt = [3]

    def new():
        nxt[0] += 1
        return V.new(nxt[0] - 1)
    st = []
    # ConditionEq on x - y: b = (x != y), inv
    b, inv = new(), new()
    diff = lc((x, 1), (y, r - 1))
    st.append(Directive([q(diff)], [b, inv], "ConditionEq"))
    st.append(Constraint(QuadComb(diff, LinComb.from_var(inv)), LinComb.from_var(b)))
    st.append(Constraint(QuadComb(lc((V.one(), 1), (b, r - 1)), diff), LinComb.zero()))
    # Bits(width) of x (big-endian), booleanity, recomposition
    bits = [new() for _ in range(width)]
    st.append(Directive([q(x)], bits, "Bits", width))
    for t in bits:
        st.append(Constraint(QuadComb(LinComb.from_var(t), LinComb.from_var(t)), LinComb.from_var(t)))
    st.append(Constraint(QuadComb(lc(*[(t, 1 << (width - 1 - i)) for i, t in enumerate(bits)]), LinComb.one()), LinComb.from_var(x)))
    b0, b1, b2 = bits[-1], bits[-2], bits[-3]
    # Xor / Or / ShaAndXorAndXorAnd / ShaCh: a directive defines the value, a constraint checks it
    xo = new(); st.append(Directive([q(b0), q(b1)], [xo], "Xor"))
    st.append(Constraint(QuadComb(lc((b0, 2)), LinComb.from_var(b1)), lc((b0, 1), (b1, 1), (xo, r - 1))))
    orr = new(); st.append(Directive([q(b0), q(b1)], [orr], "Or"))
    st.append(Constraint(QuadComb(LinComb.from_var(b0), LinComb.from_var(b1)), lc((b0, 1), (b1, 1), (orr, r - 1))))
    maj = new(); st.append(Directive([q(b0), q(b1), q(b2)], [maj], "ShaAndXorAndXorAnd"))
    ch = new(); st.append(Directive([q(b0), q(b1), q(b2)], [ch], "ShaCh"))
    st.append(Constraint(QuadComb(LinComb.from_var(b0), lc((b1, 1), (b2, r - 1))), lc((ch, 1), (b2, r - 1))))
    # Div: d = x / (y + 1) with a quadratic input, checked by d * (y + 1) = x
    dv = new(); st.append(Directive([q(x), q(lc((y, 1), (V.one(), 1)))], [dv], "Div"))
    st.append(Constraint(QuadComb(LinComb.from_var(dv), lc((y, 1), (V.one(), 1))), LinComb.from_var(x)))
    # EuclideanDiv: x = qq * (y + 1) + rr
    qq, rr = new(), new()
    st.append(Directive([q(x), q(lc((y, 1), (V.one(), 1)))], [qq, rr], "EuclideanDiv"))
    st.append(Constraint(QuadComb(LinComb.from_var(qq), lc((y, 1), (V.one(), 1))), lc((x, 1), (rr, r - 1))))
    # a variable only a directive defines (no R1CS column), a product input, and the outputs
    ghost = new(); st.append(Directive([QuadComb(LinComb.from_var(x), LinComb.from_var(y))], [ghost], "Xor") if False else
                             Directive([QuadComb(LinComb.from_var(x), LinComb.from_var(y)), q(p)], [ghost], "Or"))
    st.append(Constraint(QuadComb(lc((maj, 1), (xo, 3)), lc((orr, 1), (V.one(), 7))), LinComb.from_var(V.public(0))))
    st.append(Constraint(QuadComb(lc((qq, 1), (rr, 1), (dv, 1)), LinComb.from_var(p)), LinComb.from_var(V.public(1))))
    return Prog([Parameter.private_(x), Parameter.private_(y), Parameter.public(p)], 2, st, curve)


def random_program(curve, seed, n=60):
    """Random DAG of definitions, checks and directives over previously defined variables."""
    rnd = random.Random(seed)
    r = get_curve(curve).r
    args = [V.new(i) for i in range(3)]
    avail = list(args) + [V.one()]
    nxt = 3
    st = []

    def rlc():
        return LinComb([(rnd.choice(avail), rnd.choice([1, 2, r - 1, rnd.randrange(r)])) for _ in range(rnd.choice([1, 1, 2, 3]))])
    for j in range(n):
        kind = rnd.choice(["def", "def", "def", "check", "Bits", "Xor", "Or", "ShaCh", "ShaAndXorAndXorAnd", "Div", "ConditionEq", "EuclideanDiv"])
        if kind == "def":
            out = V.new(nxt); nxt += 1
            st.append(Constraint(QuadComb(rlc(), rlc()), LinComb.from_var(out)))
            avail.append(out)
        elif kind == "check":
            l, rr_ = rlc(), rlc()
            t = V.new(nxt); nxt += 1
            st.append(Constraint(QuadComb(l, rr_), LinComb.from_var(t)))      # defines t ...
            st.append(Constraint(QuadComb(l, rr_), lc((t, 1))))               # ... then checks the same relation
            avail.append(t)
        else:
            n_in, n_out, arg = {"Bits": (1, 0, rnd.choice([1, 5, 64, 254, 256, 300])), "Xor": (2, 1, None), "Or": (2, 1, None),
                                "ShaCh": (3, 1, None), "ShaAndXorAndXorAnd": (3, 1, None), "Div": (2, 1, None),
                                "ConditionEq": (1, 2, None), "EuclideanDiv": (2, 2, None)}[kind]
            if kind == "Bits":
                n_out = arg
            outs = [V.new(nxt + i) for i in range(n_out)]
            nxt += n_out
            st.append(Directive([QuadComb(rlc(), rlc()) for _ in range(n_in)], outs, kind, arg))
            avail.extend(outs[:4])
    out = V.public(0)
    st.append(Constraint(QuadComb(rlc(), rlc()), LinComb.from_var(out)))
    prog = Prog([Parameter.private_(args[0]), Parameter.public(args[1]), Parameter.private_(args[2])], 1, st, curve)
    return prog, [rnd.randrange(r) for _ in range(3)]


def run_native(ctx, prog, inputs, try_oor=False):
    h = ctx.prog_load(zir.write_prog(prog))
    try:
        return h, ctx.prog_info(h), ctx.prog_compute_witness(h, inputs, try_oor)
    except Exception:
        ctx.prog_free(h)
        raise


def check_program(lib, prog, inputs, try_oor=False):
    cid = 0 if prog.curve == "bn128" else 1
    ctx = Context(cid, 0, lib)
    ref = ir.Interpreter(try_oor).execute(prog, inputs)
    h, info, wit = run_native(ctx, prog, inputs, try_oor)
    try:
        assert wit == ref.write()                                         # the witness FILE, byte for byte
        r1 = synthesize(prog)
        assert (info["constraints"], info["instance"], info["witness"]) == (r1.num_constraints, r1.num_instance, r1.num_witness)
        assert info["arguments"] == len(prog.arguments) and info["returns"] == prog.return_count
        assert info["directives"] == sum(isinstance(s, Directive) for s in prog.statements)
        assert ctx.prog_public_inputs(h) == prog.public_inputs_values(ref)
        # the R1CS the program owns is the one the Python synthesis builds: same satisfied system, and the witness file
        # read back gives the same public inputs
        ctx.prog_set_witness(h, wit)
        assert ctx.prog_public_inputs(h) == prog.public_inputs_values(ref)
        assert ctx.r1cs_check(info["r1cs"], r1.assignment(ref)) is None
    finally:
        ctx.prog_free(h)
    return info


def _cases():
    yield solver_program("bn128", 8), [201, 77, 5]
    yield solver_program("bn128", 254), [2 ** 200 + 12345, 99, 3]
    yield solver_program("bls12_381", 16), [40000, 40000, 9]               # x == y: ConditionEq zero branch
    x, y = V.new(0), V.new(1)
    yield Prog([Parameter.private_(x), Parameter.public(y)], 0, [ir.constraint(x, x, y)], "bn128"), [337, 113569]
    yield Prog([], 0, [], "bn128"), []


@pytest.mark.parametrize("case", range(5))
def test_native_witness_matches_interpreter_emu(case, emu_lib):
    prog, inputs = list(_cases())[case]
    info = check_program(emu_lib, prog, inputs)
    if case == 0:
        assert info["extra_variables"] == 1 and info["unsupported_directives"] == 0 and info["schedulable"] == 1


@pytest.mark.parametrize("curve", ["bn128", "bls12_381"])
def test_native_random_programs_emu(curve, emu_lib):
    for seed in range(6):
        prog, inputs = random_program(curve, seed)
        check_program(emu_lib, prog, inputs)


@pytest.mark.parametrize("curve", ["bn128", "bls12_381"])
def test_out_of_range_bits_emu(curve, emu_lib):
    """`try_out_of_range`: x + r is decomposed instead of x when it fits the field's bit length (lib.rs:140-165)."""
    c = get_curve(curve)
    req = c.r.bit_length()
    x = V.new(0)
    for width in (req, req + 3):
        bits = [V.new(1 + i) for i in range(width)]
        prog = Prog([Parameter.private_(x)], 0, [Directive([q(x)], bits, "Bits", width)] +
                    [Constraint(QuadComb(lc(*[(t, pow(2, width - 1 - i, c.r)) for i, t in enumerate(bits)]), LinComb.one()),
                                LinComb.from_var(x))], curve)
        for val in (5, (1 << req) - c.r - 1, (1 << req) - c.r, c.r - 1):
            ref_plain = ir.Interpreter().execute(prog, [val])
            ref_oor = ir.Interpreter(True).execute(prog, [val])
            if val + c.r < (1 << req):
                assert ref_plain.values != ref_oor.values
            check_program(emu_lib, prog, [val], try_oor=True)
            check_program(emu_lib, prog, [val], try_oor=False)


def test_native_errors_emu(emu_lib):
    ctx = Context(0, 0, emu_lib)
    x, y, t = V.new(0), V.new(1), V.new(2)
    prog = Prog([Parameter.private_(x), Parameter.public(y)], 0, [ir.constraint(x, x, y)], "bn128")
    data = zir.write_prog(prog)
    h = ctx.prog_load(data)
    with pytest.raises(ZkbError, match="UNSAT|not satisfied"):
        ctx.prog_compute_witness(h, [3, 10])
    with pytest.raises(ZkbError, match="WrongInputCount"):
        ctx.prog_compute_witness(h, [3])
    with pytest.raises(ZkbError, match="witness file"):
        ctx.prog_set_witness(h, b"\x01\x00\x00\x00\x00\x00\x00\x00")
    w = Witness({V.one(): 1, x: 3}, "bn128")                                # y missing
    with pytest.raises(ZkbError, match="has no value"):
        ctx.prog_set_witness(h, w.write())
    ctx.prog_free(h)
    with pytest.raises(ZkbError):
        ctx.prog_info(h)
    for bad in (data[:50], b"XXXX" + data[4:], data[:4] + b"\x02" + data[5:], data[:-3]):
        with pytest.raises(ZkbError):
            ctx.prog_load(bad)
    with pytest.raises(ZkbError, match="another curve"):
        Context(1, 0, emu_lib).prog_load(data)
    # a read before any definition fails like the reference's unwrap (the program still loads: proving from a witness file works)
    bad_prog = Prog([Parameter.private_(x)], 0, [ir.constraint(t, x, y)], "bn128")
    h = ctx.prog_load(zir.write_prog(bad_prog))
    assert ctx.prog_info(h)["schedulable"] == 0
    with pytest.raises(ZkbError, match="no value yet"):
        ctx.prog_compute_witness(h, [3])
    ctx.prog_free(h)
    # a directive must match its solver's signature (Solver::get_signature): Xor takes two inputs
    with pytest.raises(ZkbError, match="signature"):
        ctx.prog_load(zir.write_prog(Prog([Parameter.private_(x)], 0, [Directive([q(x)], [y], "Xor")], "bn128")))
    # a Zir solver has no device path
    zprog = Prog([Parameter.private_(x)], 0, [Directive([q(x)], [y], "Zir", None)], "bn128")
    h = ctx.prog_load(zir.write_prog(zprog).replace(b"cZir", b"cZir"))
    info = ctx.prog_info(h)
    if info["unsupported_directives"]:
        with pytest.raises(ZkbError, match="no device path"):
            ctx.prog_compute_witness(h, [3])
    ctx.prog_free(h)


@pytest.mark.gpu
@pytest.mark.parametrize("case", range(3))
def test_native_witness_matches_interpreter_gpu(case, gpu_lib):
    prog, inputs = list(_cases())[case]
    check_program(gpu_lib, prog, inputs)
    prog, inputs = random_program(prog.curve, 100 + case, n=200)
    check_program(gpu_lib, prog, inputs, try_oor=bool(case & 1))


def test_program_cache_by_content_emu(emu_lib):
    """The static trait method hands the program over on every call: a second zkb_prog_load of the same bytes is a handle onto the
    resident program (also after the last handle was released), other bytes are not."""
    ctx = Context(0, 0, emu_lib)
    prog, inputs = solver_program("bn128", 8), [201, 77, 5]
    data = zir.write_prog(prog)
    h1 = ctx.prog_load(data)
    assert "prog_cache_hit" not in ctx.timings()
    h2 = ctx.prog_load(data)
    assert h2 != h1 and "prog_cache_hit" in ctx.timings()
    assert ctx.prog_info(h1)["r1cs"] == ctx.prog_info(h2)["r1cs"]
    ref = ir.Interpreter().execute(prog, inputs).write()
    assert ctx.prog_compute_witness(h2, inputs) == ref
    ctx.prog_free(h1); ctx.prog_free(h2)
    with pytest.raises(ZkbError):
        ctx.prog_info(h2)
    h3 = ctx.prog_load(data)                                                 # revived from the idle slot
    assert "prog_cache_hit" in ctx.timings()
    assert ctx.prog_compute_witness(h3, inputs) == ref
    other = zir.write_prog(solver_program("bn128", 9))
    h4 = ctx.prog_load(other)
    assert "prog_cache_hit" not in ctx.timings()
    ctx.prog_free(h3); ctx.prog_free(h4)
