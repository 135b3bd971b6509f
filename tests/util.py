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


def reference_edge_programs():
    """The program shapes of the reference's backend unit tests (zokrates_bellman/src/lib.rs:236-474: empty, identity,
    public identity, no arguments, unordered variables, `+ one`, and zokrates_ark/src/groth16.rs:125-135) as
    (name, product Prog, oracle Prog, inputs)."""
    from zokrates_b200.ir import LinComb, Parameter, Prog, Variable, constraint
    V = Variable
    one = LinComb.one()
    return [
        ("empty", Prog([], 0, [], "bn128"), oir.Prog([], 0, []), []),
        ("identity", Prog([Parameter.private_(V.new(0))], 0, [constraint(V.new(0), one, V.new(0))], "bn128"),
         oir.Prog([(1, True)], 0, [oir.Constraint([(1, 1)], [(0, 1)], [(1, 1)])]), [5]),
        ("public_identity", Prog([Parameter.public(V.new(0))], 0, [constraint(V.new(0), one, V.new(0))], "bn128"),
         oir.Prog([(1, False)], 0, [oir.Constraint([(1, 1)], [(0, 1)], [(1, 1)])]), [5]),
        ("no_arguments", Prog([], 0, [constraint(one, one, one)], "bn128"),
         oir.Prog([], 0, [oir.Constraint([(0, 1)], [(0, 1)], [(0, 1)])]), []),
        ("with_one", Prog([Parameter.private_(V.new(3))], 1,
                          [constraint(LinComb([(V.new(3), 1), (V.one(), 1)]), one, V.public(0))], "bn128"),
         oir.Prog([(4, True)], 1, [oir.Constraint([(4, 1), (0, 1)], [(0, 1)], [(-1, 1)])]), [3]),
        ("unordered_variables", Prog([Parameter.private_(V.new(42)), Parameter.public(V.new(51))], 0,
                                     [constraint(LinComb([(V.new(42), 1), (V.new(51), 1)]), one, LinComb([(V.new(7), 1)])),
                                      constraint(V.new(7), V.new(42), V.new(3))], "bn128"),
         oir.Prog([(43, True), (52, False)], 0, [oir.Constraint([(43, 1), (52, 1)], [(0, 1)], [(8, 1)]),
                                                  oir.Constraint([(8, 1)], [(43, 1)], [(4, 1)])]), [3, 4]),
        ("public_output", Prog([Parameter.public(V.new(0))], 1, [constraint(V.new(0), one, V.public(0))], "bn128"),
         oir.Prog([(1, False)], 1, [oir.Constraint([(1, 1)], [(0, 1)], [(-1, 1)])]), [42]),
    ]


def check_backend_roundtrip(lib, name, pprog, oprog, inputs):
    """setup -> generate_proof through the host mirror; compare with the oracle (key bytes, JSON, pairing)."""
    import io
    from oracle.ff import BN254
    from zokrates_b200 import backend, rng
    from zokrates_b200.ir import Interpreter
    c = BN254
    pw = Interpreter().execute(pprog, inputs)
    ow = oir.execute(c, oprog, inputs)
    td = [3, 5, 7, 11, 13, 17, 19]
    kp = backend.B200.setup(pprog, td, lib=lib)
    r1o, z = ark.synthesize(oprog, ow)
    assert kp.pk == ark.pk_serialize(c, ark.setup(c, r1o, ark.Trapdoor(*td))), name
    proof = backend.B200.generate_proof(pprog, pw, io.BytesIO(kp.pk), rng.get_rng_from_entropy("edge"), lib=lib)
    orng = ark.rng_from_entropy("edge")
    r, s = ark.fr_rand(c, orng), ark.fr_rand(c, orng)
    exp = ark.trapdoor_expected_proof(c, r1o, ark.Trapdoor(*td), z, r, s)
    pub = oprog.public_inputs_values(ow)
    assert proof.to_tagged_json() == ark.tagged_proof_json(c, exp, pub), name
    assert ark.verify(c, ark.pk_deserialize(c, kp.pk), pub, exp), name
