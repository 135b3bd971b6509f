"""Property tests (hypothesis) of the host-side formats and the witness levelizer.  CPU only, no engine needed except where noted."""
import io

import numpy as np
from hypothesis import HealthCheck, given, settings, strategies as st

from zokrates_b200 import ir, witness_gpu, zir
from zokrates_b200.curves import BN128, BLS12_381
from zokrates_b200.ir import Constraint, Directive, LinComb, Parameter, Prog, QuadComb, Variable
from zokrates_b200.r1cs import synthesize

R = BN128.r
var_ids = st.integers(min_value=-4, max_value=40)
coeffs = st.one_of(st.sampled_from([0, 1, 2, R - 1]), st.integers(min_value=0, max_value=R - 1))
lincombs = st.lists(st.tuples(var_ids.map(Variable), coeffs), max_size=4).map(LinComb)
constraints = st.builds(lambda l, r, c, e: Constraint(QuadComb(l, r), c, e), lincombs, lincombs, lincombs,
                        st.sampled_from([None, "Bitness", "Sum", "SourceAssertion"]))
directives = st.builds(lambda ins, outs, sol: Directive([QuadComb(a, b) for a, b in ins], outs, sol[0], sol[1]),
                       st.lists(st.tuples(lincombs, lincombs), max_size=3), st.lists(var_ids.map(Variable), max_size=3),
                       st.sampled_from([("Xor", None), ("Or", None), ("ConditionEq", None), ("Bits", 8), ("Bits", 254), ("ShaCh", None), ("Div", None)]))


@settings(max_examples=60, deadline=None)
@given(st.lists(st.tuples(st.integers(min_value=1, max_value=40).map(Variable), st.booleans()), max_size=4, unique_by=lambda t: t[0]),
       st.integers(min_value=0, max_value=3), st.lists(st.one_of(constraints, directives), max_size=12), st.sampled_from(["bn128", "bls12_381"]))
def test_zir_program_file_round_trip(args, n_ret, stmts, curve):
    r = BN128.r if curve == "bn128" else BLS12_381.r
    for s in stmts:                      # coefficients were drawn below the BN254 modulus, which is the smaller one
        assert all(c < r for lc in ([s.quad.left, s.quad.right, s.lin] if isinstance(s, Constraint) else []) for _, c in lc.value)
    prog = Prog([Parameter(v, p) for v, p in args], n_ret, stmts, curve)
    data = zir.write_prog(prog)
    back = zir.read_prog(data)
    assert back.curve == curve and back.arguments == prog.arguments and back.return_count == n_ret
    assert len(back.statements) == len(stmts)
    for a, b in zip(back.statements, stmts):
        assert type(a) is type(b)
        if isinstance(a, Constraint):
            assert (a.quad, a.lin, a.error) == (b.quad, b.lin, b.error)
        else:
            assert (a.solver, a.arg, a.inputs, a.outputs) == (b.solver, b.arg, b.inputs, b.outputs)
    assert zir.write_prog(back) == data                              # canonical: writing what was read gives the same bytes
    name, n_cons, n_ret2, sec = zir.read_header(data)
    assert n_cons == prog.constraint_count() and n_ret2 == n_ret and sec[3][1] + sec[3][2] == len(data)


@settings(max_examples=40, deadline=None)
@given(st.dictionaries(var_ids.map(Variable), st.integers(min_value=0, max_value=R - 1), max_size=12))
def test_witness_file_round_trip(values):
    w = ir.Witness(values)
    data = w.write()
    assert ir.Witness.read(data).values == w.values and len(data) == 8 + len(values) * 40
    ids = [v.id for v, _ in w.items()]
    assert ids == sorted(ids)                                        # BTreeMap order: ascending signed id


@st.composite
def dag_programs(draw):
    """Directive-free programs whose constraints assign fresh variables from earlier ones (plus a few consistent checks)."""
    n_in = draw(st.integers(min_value=1, max_value=3))
    n = draw(st.integers(min_value=1, max_value=14))
    args = [Parameter(Variable.new(i), draw(st.booleans())) for i in range(n_in)]
    defined = [Variable.one()] + [a.id for a in args]
    stmts = []
    lc = lambda: LinComb([(draw(st.sampled_from(defined)), draw(coeffs)) for _ in range(draw(st.integers(min_value=1, max_value=3)))])
    for k in range(n):
        out = Variable.new(n_in + k)
        q = QuadComb(lc(), lc())
        stmts.append(Constraint(q, LinComb.from_var(out)))
        defined.append(out)
        if draw(st.booleans()):                                       # a check that holds: the same product against the new variable
            stmts.append(Constraint(q, LinComb([(out, 1), (Variable.one(), 0)])))
    inputs = [draw(st.integers(min_value=0, max_value=R - 1)) for _ in range(n_in)]
    return Prog(args, 0, stmts), inputs


@settings(max_examples=25, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
@given(dag_programs())
def test_levelised_evaluation_equals_the_interpreter(emu_lib, case):
    prog, inputs = case
    ref = ir.Interpreter().execute(prog, inputs)
    got = witness_gpu.generate_witness(prog, inputs, lib=emu_lib)
    assert got.values == ref.values
    r1cs = synthesize(prog)
    cols = {v: i for i, v in enumerate(r1cs.instance_vars)}
    cols.update({v: r1cs.num_instance + i for i, v in enumerate(r1cs.witness_vars)})
    level_ptr, rows, out_var = witness_gpu.levelize(r1cs, [0] + [cols[p.id] for p in prog.arguments])
    assert level_ptr[-1] == r1cs.num_constraints and sorted(rows.tolist()) == list(range(r1cs.num_constraints))
    assert np.all(np.diff(level_ptr.astype(np.int64)) > 0)             # no empty level
    fast = witness_gpu.levelize_wavefront(r1cs, [0] + [cols[p.id] for p in prog.arguments])
    assert fast is not None and all(np.array_equal(x, y) for x, y in zip(fast, (level_ptr, rows, out_var)))
