"""Shared helpers for the parity tests (test infrastructure)."""
import random

import numpy as np

from oracle import ark, ir as oir
from zokrates_b200 import ir as pir
from zokrates_b200._lib import fr_array


def rand_prog_pair(c, ncons, npub, npriv, seed, curve_name):
    """The same random program as (oracle Prog, product Prog)."""
    rnd = random.Random(seed)
    oargs = [(oir.var_new(i), i >= npub) for i in range(npub + npriv)]
    nxt = npub + npriv
    ostmts, pstmts = [], []
    avail = [v for v, _ in oargs]
    for j in range(ncons):
        def lc():
            return [(rnd.choice(avail + [0]), rnd.choice([1, 1, 2, c.r - 1, rnd.randrange(c.r)]))
                    for _ in range(rnd.choice([1, 1, 2, 3]))]
        out = oir.var_new(nxt)
        nxt += 1
        if j == ncons - 1:
            out = oir.var_public(0)
        l, r_, o = lc(), lc(), [(out, 1)]
        ostmts.append(oir.Constraint(l, r_, o))

        def plc(terms):
            return pir.LinComb([(pir.Variable(v), k) for v, k in terms])
        pstmts.append(pir.Constraint(pir.QuadComb(plc(l), plc(r_)), plc(o)))
        avail.append(out)
    oprog = oir.Prog(oargs, 1, ostmts)
    pprog = pir.Prog([pir.Parameter(pir.Variable(v), priv) for v, priv in oargs], 1, pstmts, curve_name)
    inputs = [rnd.randrange(c.r) for _ in range(npub + npriv)]
    return oprog, pprog, inputs


def proof_bytes(c, proof):
    return ark.ser_g1(c, proof[0]) + ark.ser_g2(c, proof[1]) + ark.ser_g1(c, proof[2])


def csr_from_rows(rows, r):
    rowptr = [0]
    col, val = [], []
    for row in rows:
        for cidx, k in row:
            col.append(cidx)
            val.append(k % r)
        rowptr.append(len(col))
    return np.array(rowptr, dtype=np.uint64), np.array(col, dtype=np.uint32), fr_array(val)
