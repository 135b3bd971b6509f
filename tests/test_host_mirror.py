"""Host-side mirror of the reference interface: IR, witness file, proof JSON, R1CS order.  CPU only."""
import json

import pytest

from zokrates_b200 import curves, ir, proof as pproof, r1cs
from zokrates_b200.ir import Constraint, Directive, Interpreter, LinComb, Parameter, Prog, QuadComb, Variable, Witness


def test_variable_display_and_order():
    # zokrates_ast/src/common/flat/variable.rs:82-102
    assert str(Variable.one()) == "~one" and str(Variable.public(42)) == "~out_42" and str(Variable.new(42)) == "_42"
    assert sorted([Variable.new(42), Variable.public(8), Variable.one()]) == [Variable.public(8), Variable.one(), Variable.new(42)]


def test_witness_binary_and_json():
    # zokrates_ast/src/ir/witness.rs:108-154 — exact JSON text and BTreeMap ordering
    w = Witness({Variable.new(42): 42, Variable.public(8): 8, Variable.one(): 1})
    assert Witness.read(w.write()).values == w.values
    assert w.write_json() == '{\n  "~out_8": "8",\n  "~one": "1",\n  "_42": "42"\n}'
    assert len(w.write()) == 8 + 3 * (8 + 32)
    with pytest.raises(ValueError):
        Witness.read(w.write()[:-1])
    bad = bytearray(w.write())
    bad[-1] = 0xFF       # value >= modulus
    with pytest.raises(ValueError):
        Witness.read(bytes(bad))


def test_curve_ids():
    assert curves.BN128.field_id.hex() == "b4f7b5bd"        # zokrates_book/src/toolbox/ir.md
    assert curves.BLS12_381.field_id.hex() == "40d8c1f9"


def test_interpreter_and_public_inputs():
    # the program of zokrates_ark/src/groth16.rs:125-135: public x, one return, x * 1 == ~out_0... here `_0 == ~out_0`
    prog = Prog([Parameter.public(Variable.new(0))], 1, [ir.constraint(Variable.new(0), LinComb.one(), Variable.public(0))])
    w = Interpreter().execute(prog, [42])
    assert w[Variable.public(0)] == 42 and w[Variable.one()] == 1
    assert prog.public_inputs_values(w) == [42, 42]
    with pytest.raises(ValueError):
        Interpreter().execute(prog, [1, 2])


def test_solvers():
    # zokrates_interpreter/src/lib.rs:426-511
    c = curves.BN128
    assert Interpreter.execute_solver(c, "ConditionEq", None, [0]) == [0, 1]
    assert Interpreter.execute_solver(c, "ConditionEq", None, [1]) == [1, 1]
    res = Interpreter.execute_solver(c, "Bits", 254, [42])
    assert res[247:] == [0, 1, 0, 1, 0, 1, 0] and not any(res[:248])
    assert Interpreter.execute_solver(c, "Bits", 500, [1]) == [0] * 499 + [1]
    assert Interpreter.execute_solver(c, "Xor", None, [1, 1]) == [0]
    assert Interpreter.execute_solver(c, "EuclideanDiv", None, [17, 5]) == [3, 2]
    assert Interpreter.execute_solver(c, "Div", None, [6, 0]) == [1]


def test_unsatisfied_and_directive():
    x, y, b0 = Variable.new(0), Variable.new(1), Variable.new(2)
    prog = Prog([Parameter.private_(x)], 0, [
        Directive([QuadComb(LinComb.one(), LinComb.from_var(x))], [y, b0], "ConditionEq"),
        ir.constraint(x, b0, y),     # x * (1/x) == 1
    ])
    w = Interpreter().execute(prog, [5])
    assert w[y] == 1 and w[b0] == pow(5, -1, curves.BN128.r)
    bad = Prog([Parameter.private_(x), Parameter.public(y)], 0, [ir.constraint(x, x, y), ir.constraint(x, LinComb.one(), y)])
    with pytest.raises(ir.UnsatisfiedConstraint):
        Interpreter().execute(bad, [3, 9])


def test_ark_variable_order():
    """zokrates_ark/src/lib.rs:80-130: one, public args / outputs as instance in allocation order; private
    args then first appearance (left, right, lin) as witness; unordered ids keep first-appearance order."""
    a, b, c_, d = Variable.new(0), Variable.new(1), Variable.new(5), Variable.new(3)
    prog = Prog([Parameter.private_(a), Parameter.public(b)], 1, [
        ir.constraint(c_, d, Variable.public(0)),            # c_ and d appear before any smaller id
        ir.constraint(a, LinComb([(b, 2), (b, 3)]), d),       # duplicate terms are kept
    ])
    r = r1cs.synthesize(prog)
    assert r.instance_vars == [Variable.one(), b, Variable.public(0)]
    assert r.witness_vars == [a, c_, d]
    assert (r.num_instance, r.num_witness, r.num_constraints, r.domain_size) == (3, 3, 2, 8)
    assert list(r.a[1]) == [4, 3] and list(r.b[1]) == [5, 1, 1] and list(r.c[1]) == [2, 5]
    assert list(r.b[0]) == [0, 1, 3]


def test_proof_json_shape_and_roundtrip():
    c = curves.BN128
    raw = b"".join(int(i + 1).to_bytes(32, "little") for i in range(8))
    p = pproof.Proof.from_raw(c, raw, [5, 113569])
    d = json.loads(p.to_tagged_json())
    assert list(d.keys()) == ["scheme", "curve", "proof", "inputs"] and d["scheme"] == "g16" and d["curve"] == "bn128"
    assert d["proof"]["a"] == ["0x" + "00" * 31 + "01", "0x" + "00" * 31 + "02"]
    assert d["proof"]["b"] == [["0x" + "00" * 31 + "03", "0x" + "00" * 31 + "04"], ["0x" + "00" * 31 + "05", "0x" + "00" * 31 + "06"]]
    assert d["inputs"][1] == "0x" + (113569).to_bytes(32, "big").hex()
    assert pproof.Proof.from_json(p.to_tagged_json()).to_raw() == raw
    assert p.input_values() == [5, 113569]
    p2 = pproof.Proof.from_raw(curves.BLS12_381, b"".join(int(i + 1).to_bytes(48, "little") for i in range(8)), [])
    assert len(json.loads(p2.to_tagged_json())["proof"]["a"][0]) == 2 + 96
