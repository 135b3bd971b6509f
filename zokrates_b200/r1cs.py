"""IR program -> R1CS matrices in ark-relations order, as CSR arrays for `zkb_r1cs_load`.

Mirrors `Computation::generate_constraints` + `ark_combination`
(/root/reference/zokrates_ark/src/lib.rs:41-130): index 0 is the constant one; public arguments and
outputs (`id < 0`) become instance variables, everything else witness variables, each numbered in
allocation order (arguments first, then first appearance scanning quad.left, quad.right, lin of every
constraint).  Matrix column of witness j is num_instance + j.  Duplicate variables inside one
combination are kept as separate terms (the reference does not merge them either; the sums agree).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Tuple

import numpy as np

from ._lib import fr_array
from .curves import curve as _curve
from .ir import Constraint, Prog, Variable, Witness


@dataclass
class R1CS:
    curve: str
    num_constraints: int
    num_instance: int          # incl. the constant one
    num_witness: int
    a: Tuple[np.ndarray, np.ndarray, np.ndarray]   # rowptr uint64[N+1], col uint32[nnz], val uint64[nnz,4]
    b: Tuple[np.ndarray, np.ndarray, np.ndarray]
    c: Tuple[np.ndarray, np.ndarray, np.ndarray]
    instance_vars: Optional[List[Variable]] = None  # IR variable of every instance column (None for raw CSR input)
    witness_vars: Optional[List[Variable]] = None

    @property
    def num_variables(self) -> int:
        return self.num_instance + self.num_witness

    @property
    def domain_size(self) -> int:
        n = 1
        while n < self.num_constraints + self.num_instance:
            n <<= 1
        return n

    def matrices(self):
        return [self.a, self.b, self.c]

    def assignment(self, witness: Witness) -> np.ndarray:
        """Full assignment z = [1, instance.., witness..] as canonical LE limbs (uint64[m,4])."""
        vals = [witness[v] for v in self.instance_vars] + [witness[v] for v in self.witness_vars]
        return fr_array(vals)


def synthesize(prog: Prog) -> R1CS:
    c = _curve(prog.curve)
    symbols = {Variable.one(): ("i", 0)}
    inst: List[Variable] = [Variable.one()]
    wit: List[Variable] = []
    for p in prog.arguments:
        if p.id in symbols:
            raise ValueError("duplicate argument")
        if p.private:
            symbols[p.id] = ("w", len(wit)); wit.append(p.id)
        else:
            symbols[p.id] = ("i", len(inst)); inst.append(p.id)

    rows = ([], [], [])   # per matrix: list of term lists

    def comb(lc):
        out = []
        for v, coeff in lc.value:
            sym = symbols.get(v)
            if sym is None:
                if v.is_output():
                    sym = ("i", len(inst)); inst.append(v)
                else:
                    sym = ("w", len(wit)); wit.append(v)
                symbols[v] = sym
            out.append((sym, coeff % c.r))
        return out

    for s in prog.statements:
        if isinstance(s, Constraint):
            rows[0].append(comb(s.quad.left))
            rows[1].append(comb(s.quad.right))
            rows[2].append(comb(s.lin))

    ni = len(inst)

    def csr(terms_rows):
        rowptr = np.zeros(len(terms_rows) + 1, dtype=np.uint64)
        cols, vals = [], []
        for i, row in enumerate(terms_rows):
            for (kind, idx), coeff in row:
                cols.append(idx if kind == "i" else ni + idx)
                vals.append(coeff)
            rowptr[i + 1] = len(cols)
        return rowptr, np.array(cols, dtype=np.uint32), fr_array(vals)

    return R1CS(c.name, len(rows[0]), ni, len(wit), csr(rows[0]), csr(rows[1]), csr(rows[2]), inst, wit)
