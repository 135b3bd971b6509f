"""Host-side mirror of the ZoKrates IR data model the proving backend consumes.

Mirrors, with the same names and semantics (all paths under /root/reference):
  * `Variable`            zokrates_ast/src/common/flat/variable.rs:6-60
  * `LinComb`/`QuadComb`  zokrates_ast/src/ir/expression.rs:10-18,72-78
  * `Statement`           zokrates_ast/src/ir/mod.rs:118-128  (Constraint / Directive / Log)
  * `Prog`                zokrates_ast/src/ir/mod.rs:211-288  (`ProgIterator`)
  * `Witness`             zokrates_ast/src/ir/witness.rs:8-83 (binary + JSON forms, BTreeMap order)
  * `Interpreter`         zokrates_interpreter/src/lib.rs:40-138,249-307,366-378
Field elements are Python ints reduced mod the curve's scalar field.
"""
from __future__ import annotations

import io
import json
import struct
from dataclasses import dataclass, field
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

from .curves import Curve, curve as _curve


@dataclass(frozen=True, order=True)
class Variable:
    """id > 0 intermediate, id == 0 `~one`, id < 0 public outputs (variable.rs:6-12)."""
    id: int

    @staticmethod
    def new(i: int) -> "Variable":
        return Variable(1 + i)

    @staticmethod
    def one() -> "Variable":
        return Variable(0)

    @staticmethod
    def public(i: int) -> "Variable":
        return Variable(-i - 1)

    def is_output(self) -> bool:
        return self.id < 0

    def __str__(self):
        if self.id == 0:
            return "~one"
        return f"_{self.id - 1}" if self.id > 0 else f"~out_{-(self.id + 1)}"

    def write(self) -> bytes:
        return struct.pack("<q", self.id)


@dataclass
class Parameter:
    id: Variable
    private: bool

    @staticmethod
    def public(v: Variable) -> "Parameter":
        return Parameter(v, False)

    @staticmethod
    def private_(v: Variable) -> "Parameter":
        return Parameter(v, True)


@dataclass
class LinComb:
    """Sum of coefficient * variable terms; duplicates are kept (expression.rs:72-78)."""
    value: List[Tuple[Variable, int]] = field(default_factory=list)

    @staticmethod
    def summand(coeff: int, v: Variable) -> "LinComb":
        return LinComb([(v, coeff)])

    @staticmethod
    def from_var(v: Variable) -> "LinComb":
        return LinComb([(v, 1)])

    @staticmethod
    def one() -> "LinComb":
        return LinComb([(Variable.one(), 1)])

    @staticmethod
    def zero() -> "LinComb":
        return LinComb([])

    def __add__(self, other: "LinComb") -> "LinComb":
        return LinComb(self.value + other.value)


@dataclass
class QuadComb:
    left: LinComb
    right: LinComb


@dataclass
class Constraint:
    """`quad.left * quad.right == lin`  (ConstraintStatement, ir/mod.rs:34-43)."""
    quad: QuadComb
    lin: LinComb
    error: Optional[str] = None


@dataclass
class Directive:
    """outputs = solver(inputs); skipped by the proving backend (zokrates_ark/src/lib.rs:116)."""
    inputs: List[QuadComb]
    outputs: List[Variable]
    solver: str
    arg: Optional[int] = None


@dataclass
class Log:
    format_string: str
    expressions: list = field(default_factory=list)


def constraint(left, right, lin) -> Constraint:
    """`Statement::constraint(quad, lin, error)` helper accepting Variables / LinCombs."""
    def lc(x):
        return x if isinstance(x, LinComb) else LinComb.from_var(x)
    return Constraint(QuadComb(lc(left), lc(right)), lc(lin))


def definition(v: Variable, quad_or_lin) -> Constraint:
    """`Statement::definition(v, e)`: e == v."""
    if isinstance(quad_or_lin, QuadComb):
        return Constraint(quad_or_lin, LinComb.from_var(v))
    lin = quad_or_lin if isinstance(quad_or_lin, LinComb) else LinComb.from_var(quad_or_lin)
    return Constraint(QuadComb(LinComb.one(), lin), LinComb.from_var(v))


@dataclass
class Prog:
    arguments: List[Parameter] = field(default_factory=list)
    return_count: int = 0
    statements: list = field(default_factory=list)
    curve: str = "bn128"

    def constraint_count(self) -> int:
        return sum(isinstance(s, Constraint) for s in self.statements)

    def public_count(self) -> int:
        return sum(not a.private for a in self.arguments) + self.return_count

    def public_inputs_values(self, witness: "Witness") -> List[int]:
        """ir/mod.rs:278-288: public arguments in order, then the return values."""
        return [witness[p.id] for p in self.arguments if not p.private] + witness.return_values()


class Witness:
    """BTreeMap<Variable, T> (ir/witness.rs:8-9)."""

    def __init__(self, values: Optional[Dict[Variable, int]] = None, curve: str = "bn128"):
        self.curve = _curve(curve)
        self.values: Dict[Variable, int] = dict(values or {})

    def __getitem__(self, v: Variable) -> int:
        return self.values[v]

    def __contains__(self, v: Variable) -> bool:
        return v in self.values

    def insert(self, v: Variable, val: int):
        self.values[v] = val % self.curve.r

    def items(self):
        return sorted(self.values.items())          # BTreeMap order = ascending signed id

    def return_values(self) -> List[int]:
        """ir/witness.rs:12-24: ~out_0, ~out_1, ... ."""
        n = sum(v.is_output() for v in self.values)
        return [self.values[Variable.public(i)] for i in range(n)]

    def write(self) -> bytes:
        """ir/witness.rs:44-53: usize LE length, then (isize LE id, canonical LE value)."""
        out = io.BytesIO()
        out.write(struct.pack("<Q", len(self.values)))
        for v, val in self.items():
            out.write(v.write())
            out.write(int(val).to_bytes(self.curve.fr_bytes, "little"))
        return out.getvalue()

    @classmethod
    def read(cls, data: bytes, curve: str = "bn128") -> "Witness":
        c = _curve(curve)
        if len(data) < 8:
            raise ValueError("witness file truncated")
        (n,) = struct.unpack_from("<Q", data, 0)
        rec = 8 + c.fr_bytes
        if len(data) < 8 + n * rec:
            raise ValueError("witness file truncated")
        w = cls(curve=c)
        off = 8
        for _ in range(n):
            (vid,) = struct.unpack_from("<q", data, off)
            val = int.from_bytes(data[off + 8:off + rec], "little")
            if val >= c.r:
                raise ValueError("non-canonical field element in witness")
            w.values[Variable(vid)] = val
            off += rec
        return w

    def write_json(self) -> str:
        """ir/witness.rs:73-82."""
        return json.dumps({str(v): str(val) for v, val in self.items()}, indent=2)


class UnsatisfiedConstraint(Exception):
    pass


class Interpreter:
    """Witness generation restated from zokrates_interpreter/src/lib.rs (simple solvers only; `Zir`
    folded functions and the embed gadgets stay with the reference's compiler front end).
    `try_out_of_range` mirrors `Interpreter::try_out_of_range()` (lib.rs:32-38)."""

    def __init__(self, try_out_of_range: bool = False):
        self.should_try_out_of_range = try_out_of_range

    @staticmethod
    def try_solve_with_out_of_range_bits(c: Curve, bit_width: int, x: int) -> List[int]:
        """lib.rs:140-165: the second decomposition x + r when it still fits `get_required_bits()` bits."""
        req = c.r.bit_length()
        cand = x % c.r + c.r
        v = cand if cand < (1 << req) else x % c.r
        return [0] * (bit_width - req) + [(v >> i) & 1 for i in range(req - 1, -1, -1)]

    @staticmethod
    def evaluate_lin(c: Curve, w: Witness, l: LinComb) -> int:
        acc = 0
        for var, mult in l.value:
            acc = (acc + w.values[var] * mult) % c.r
        return acc

    @classmethod
    def evaluate_quad(cls, c: Curve, w: Witness, q: QuadComb) -> int:
        return cls.evaluate_lin(c, w, q.left) * cls.evaluate_lin(c, w, q.right) % c.r

    @staticmethod
    def execute_solver(c: Curve, solver: str, arg, x: Sequence[int]) -> List[int]:
        r = c.r
        if solver == "ConditionEq":
            return [0, 1] if x[0] % r == 0 else [1, pow(x[0], -1, r)]
        if solver == "Bits":
            v = x[0] % r
            return [(v >> (arg - 1 - i)) & 1 for i in range(arg)]
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
            return [x[0] * pow(x[1], -1, r) % r if x[1] % r else 1]
        if solver == "EuclideanDiv":
            n, d = x[0] % r, x[1] % r
            q = n // d if d else 0
            return [q, n - d * q]
        raise NotImplementedError(f"solver {solver}")

    def execute(self, prog: Prog, inputs: Sequence[int]) -> Witness:
        c = _curve(prog.curve)
        if len(inputs) != len(prog.arguments):
            raise ValueError(f"WrongInputCount: expected {len(prog.arguments)}, received {len(inputs)}")
        w = Witness(curve=c)
        w.insert(Variable.one(), 1)
        for p, val in zip(prog.arguments, inputs):
            w.insert(p.id, val)
        for s in prog.statements:
            if isinstance(s, Constraint):
                lin = s.lin.value
                is_assignee = len(lin) == 1 and lin[0][1] % c.r == 1 and lin[0][0] not in w
                q = self.evaluate_quad(c, w, s.quad)
                if is_assignee:
                    w.insert(lin[0][0], q)
                elif q != self.evaluate_lin(c, w, s.lin):
                    raise UnsatisfiedConstraint(s.error)
            elif isinstance(s, Directive):
                xs = [self.evaluate_quad(c, w, q) for q in s.inputs]
                if s.solver == "Bits" and self.should_try_out_of_range and s.arg >= c.r.bit_length():
                    res = self.try_solve_with_out_of_range_bits(c, s.arg, xs[-1])     # lib.rs:94-101
                else:
                    res = self.execute_solver(c, s.solver, s.arg, xs)
                for o, val in zip(s.outputs, res):
                    w.insert(o, val)
        return w
