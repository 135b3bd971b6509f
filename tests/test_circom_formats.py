"""`.r1cs` / `.wtns` byte layout pinned to the reference's golden tests
(zokrates_circom/src/r1cs.rs:242-430, src/witness.rs:113-230); values transcribed as test vectors."""
from zokrates_b200 import circom, ir
from zokrates_b200._lib import fr_from_array
from zokrates_b200.ir import LinComb, Parameter, Prog, QuadComb, Variable, Witness

MOD = bytes.fromhex("010000f093f5e1439170b97948e833285d588181b64550b829a031e1724e6430")


def one32(v=1):
    return int(v).to_bytes(32, "little")


def u32(v):
    return int(v).to_bytes(4, "little")


def u64(v):
    return int(v).to_bytes(8, "little")


def header(n_wires, n_out, n_pub, n_prv, n_cons):
    return u32(1) + u64(64) + u32(32) + MOD + u32(n_wires) + u32(n_out) + u32(n_pub) + u32(n_prv) + u64(n_wires) + u32(n_cons)


def test_r1cs_empty():
    exp = b"r1cs" + u32(1) + u32(3) + u32(2) + u64(0) + header(1, 0, 0, 0, 0) + u32(3) + u64(8) + u64(0)
    assert circom.write_r1cs(Prog()) == exp
    r = circom.read_r1cs(exp)
    assert (r.num_constraints, r.num_instance, r.num_witness) == (0, 1, 0)


def test_r1cs_return_one():
    prog = Prog([], 1, [ir.constraint(LinComb.one(), LinComb.one(), Variable.public(0))])
    term = lambda w: u32(1) + u32(w) + one32()
    exp = (b"r1cs" + u32(1) + u32(3) + u32(2) + u64(0x78) + term(0) + term(0) + term(1) + header(2, 1, 0, 0, 1)
           + u32(3) + u64(16) + u64(0) + u64(1))
    assert circom.write_r1cs(prog) == exp


def test_r1cs_with_inputs():
    x0, x1 = Variable.new(0), Variable.new(1)
    prog = Prog([Parameter.private_(x0), Parameter.public(x1)], 1, [
        ir.Constraint(QuadComb(LinComb.from_var(x0), LinComb.from_var(x0)), LinComb.from_var(x0)),
        ir.Constraint(QuadComb(LinComb.one(), LinComb.from_var(x0) + LinComb.from_var(x1)), LinComb.from_var(Variable.public(0))),
    ])
    t = lambda w: u32(w) + one32()
    body = (u32(1) + t(3) + u32(1) + t(3) + u32(1) + t(3)            # first constraint: wire 3 = _0
            + u32(1) + t(0) + u32(2) + t(3) + t(2) + u32(1) + t(1))  # second: one * (_0 + _1) = ~out_0
    exp = (b"r1cs" + u32(1) + u32(3) + u32(2) + u64(0x114) + body + header(4, 1, 1, 1, 2)
           + u32(3) + u64(32) + u64(0) + u64(1) + u64(2) + u64(3))
    data = circom.write_r1cs(prog)
    assert data == exp
    variables, off, _ = circom.r1cs_program(prog)
    assert variables == [Variable.one(), Variable.public(0), x1, x0] and off == 3
    r = circom.read_r1cs(data)
    assert (r.num_constraints, r.num_instance, r.num_witness, r.curve) == (2, 3, 1, "bn128")
    assert list(r.b[1]) == [3, 3, 2] and list(r.c[1]) == [3, 1]


def wt_header(n):
    return b"wtns" + u32(2) + u32(2) + u32(1) + u64(0x28) + u32(32) + MOD + u32(n)


def test_wtns_golden():
    assert circom.write_witness(Witness(), []) == wt_header(0) + u32(2) + u64(0)
    assert circom.write_witness(Witness({Variable.public(0): 1}), []) == wt_header(1) + u32(2) + u64(0x20) + one32(1)
    w = Witness({Variable.public(0): 42, Variable.one(): 1, Variable.new(0): 43, Variable.new(1): 44})
    data = circom.write_witness(w, [Variable.new(1)])
    assert data == wt_header(4) + u32(2) + u64(0x80) + one32(1) + one32(42) + one32(44) + one32(43)
    name, z = circom.read_wtns(data)
    assert name == "bn128" and fr_from_array(z) == [1, 42, 44, 43]


def test_import_and_prove(emu_lib, oracle_c):
    """export -> import -> setup + prove on the imported system (host-emulated engine) == CPU oracle bytes."""
    from zokrates_b200._lib import Context
    x0, x1, t = Variable.new(0), Variable.new(1), Variable.new(2)
    prog = Prog([Parameter.private_(x0), Parameter.public(x1)], 1, [
        ir.constraint(x0, x1, t),
        ir.constraint(LinComb([(t, 3), (x0, 5)]), LinComb([(x1, 7), (Variable.one(), 2)]), Variable.public(0)),
    ])
    w = ir.Interpreter().execute(prog, [11, 13])
    r = circom.read_r1cs(circom.write_r1cs(prog))
    name, z = circom.read_wtns(circom.write_witness(w, [p.id for p in prog.arguments if not p.private]))
    assert fr_from_array(z)[:3] == [1, w[Variable.public(0)], 13]
    ctx = Context(0, 0, emu_lib)
    h = ctx.r1cs_load(r.num_constraints, r.num_instance, r.num_witness, r.matrices())
    pk = ctx.setup(h, [3, 5, 7, 11, 13, 17, 19])
    proof = ctx.prove(ctx.pk_load(pk), h, z, 21, 22)
    ref, _ = oracle_c.prove(0, pk, r, z, 21, 22, 32)
    assert proof == ref
