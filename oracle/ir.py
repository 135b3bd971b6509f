"""ZoKrates IR data model, witness file format and interpreter semantics — oracle restatement.

Test infrastructure only (see oracle/__init__.py).  Follows, in /root/reference:
  * zokrates_ast/src/common/flat/variable.rs:6-50   Variable ids (0 = ~one, <0 outputs, >0 aux)
  * zokrates_ast/src/ir/expression.rs:10-18,72-78   QuadComb / LinComb (term lists, duplicates kept)
  * zokrates_ast/src/ir/mod.rs:118-128,211-288       Statement / ProgIterator / public_inputs_values
  * zokrates_ast/src/ir/witness.rs:8-83              Witness (BTreeMap order, binary + JSON forms)
  * zokrates_interpreter/src/lib.rs:40-138,249-307,366-378  execution, simple solvers, evaluate_lin/quad
"""
from __future__ import annotations

import io
import json
import struct
from dataclasses import dataclass, field
from typing import List, Optional, Tuple

from .ff import CurveParams, inv_mod

ONE = 0


def var_new(i: int) -> int:      # Variable::new   variable.rs:16-20
    return 1 + i


def var_public(i: int) -> int:   # Variable::public variable.rs:26-30
    return -i - 1


def var_name(v: int) -> str:     # Display          variable.rs:52-60
    if v == 0:
        return "~one"
    return f"_{v - 1}" if v > 0 else f"~out_{-(v + 1)}"


LinComb = List[Tuple[int, int]]  # [(variable id, coefficient)]


@dataclass
class Constraint:                # ConstraintStatement  ir/mod.rs:34-43 : quad.left * quad.right == lin
    left: LinComb
    right: LinComb
    lin: LinComb


@dataclass
class Directive:                 # DirectiveStatement: outputs = solver(inputs)
    inputs: List[Tuple[LinComb, LinComb]]
    outputs: List[int]
    solver: str
    arg: Optional[int] = None    # Bits(w)


@dataclass
class Prog:
    arguments: List[Tuple[int, bool]]   # (variable id, private)
    return_count: int
    statements: list = field(default_factory=list)

    def constraint_count(self):
        return sum(isinstance(s, Constraint) for s in self.statements)

    def public_inputs_values(self, witness: dict) -> List[int]:
        """ir/mod.rs:278-288: public args in declaration order, then ~out_0.. ."""
        vals = [witness[v] for v, private in self.arguments if not private]
        outs = sorted((v for v in witness if v < 0), reverse=True)  # ~out_0 = -1, ~out_1 = -2 ...
        assert outs == [var_public(i) for i in range(len(outs))]
        return vals + [witness[v] for v in outs]


# ----------------------------------------------------------------------------- witness file
def witness_write(witness: dict, fr_bytes: int = 32) -> bytes:
    """ir/witness.rs:44-53: usize LE count, then (isize LE id, canonical LE value) in BTreeMap order."""
    out = io.BytesIO()
    out.write(struct.pack("<Q", len(witness)))
    for v in sorted(witness):
        out.write(struct.pack("<q", v))
        out.write(int(witness[v]).to_bytes(fr_bytes, "little"))
    return out.getvalue()


def witness_read(data: bytes, fr_bytes: int = 32) -> dict:
    (n,) = struct.unpack_from("<Q", data, 0)
    off = 8
    w = {}
    for _ in range(n):
        (v,) = struct.unpack_from("<q", data, off)
        off += 8
        w[v] = int.from_bytes(data[off:off + fr_bytes], "little")
        off += fr_bytes
    return w


def witness_json(witness: dict) -> str:
    """ir/witness.rs:73-82 (serde_json::to_writer_pretty of name -> decimal string, BTreeMap order)."""
    return json.dumps({var_name(v): str(witness[v]) for v in sorted(witness)}, indent=2)


# ----------------------------------------------------------------------------- interpreter
class UnsatisfiedConstraint(Exception):
    pass


def evaluate_lin(c: CurveParams, w: dict, l: LinComb) -> int:
    acc = 0
    for var, mult in l:
        acc = (acc + w[var] * mult) % c.r     # KeyError == EvaluationError (lib.rs:366-372)
    return acc


def evaluate_quad(c: CurveParams, w: dict, left: LinComb, right: LinComb) -> int:
    return evaluate_lin(c, w, left) * evaluate_lin(c, w, right) % c.r


def execute_solver(c: CurveParams, solver: str, arg, x: List[int]) -> List[int]:
    r = c.r
    if solver == "ConditionEq":                                # lib.rs:249-255
        if x[0] % r == 0:
            return [0, 1]
        return [1, inv_mod(x[0], r)]
    if solver == "Bits":                                       # lib.rs:256-269 (big-endian, padded)
        v = x[0] % r
        bits = [(v >> i) & 1 for i in range(v.bit_length() - 1, -1, -1)] if v else []
        # to_bits_be yields the full repr width; keeping the `arg` least significant is equivalent
        full = [0] * max(0, r.bit_length() - len(bits)) + bits
        bits = full[max(0, len(full) - arg):]
        return [0] * (arg - len(bits)) + bits
    if solver == "Xor":
        return [(x[0] + x[1] - 2 * x[0] * x[1]) % r]
    if solver == "Or":
        return [(x[0] + x[1] - x[0] * x[1]) % r]
    if solver == "ShaAndXorAndXorAnd":
        a, b, cc = x
        return [(b * cc - (2 * b * cc - b - cc) * a) % r]
    if solver == "ShaCh":
        a, b, cc = x
        return [(a * (b - cc) + cc) % r]
    if solver == "Div":
        return [x[0] * inv_mod(x[1], r) % r if x[1] % r else 1]
    if solver == "EuclideanDiv":
        n, d = x[0] % r, x[1] % r
        q = n // d if d else 0
        return [q, n - d * q]
    raise NotImplementedError(solver)


def execute(c: CurveParams, prog: Prog, inputs: List[int]) -> dict:
    """Interpreter::execute_with_log_stream  lib.rs:40-138."""
    if len(inputs) != len(prog.arguments):
        raise ValueError("WrongInputCount")
    w = {ONE: 1}
    for (v, _), val in zip(prog.arguments, inputs):
        w[v] = val % c.r
    for s in prog.statements:
        if isinstance(s, Constraint):
            is_assignee = len(s.lin) == 1 and s.lin[0][1] % c.r == 1 and s.lin[0][0] not in w
            q = evaluate_quad(c, w, s.left, s.right)
            if is_assignee:
                w[s.lin[0][0]] = q
            elif q != evaluate_lin(c, w, s.lin):
                raise UnsatisfiedConstraint()
        elif isinstance(s, Directive):
            xs = [evaluate_quad(c, w, l, rr) for l, rr in s.inputs]
            for o, val in zip(s.outputs, execute_solver(c, s.solver, s.arg, xs)):
                w[o] = val % c.r
    return w
