"""GPU witness evaluation (`zkb_witness_eval`, `zkb_r1cs_check`) against the host mirror of the reference interpreter.
The CPU tier drives the kernel bodies through the host-emulation library; the GPU tier (-m gpu) through libzkb200.so."""
import random

import numpy as np
import pytest

from tests.util import rand_prog_pair
from zokrates_b200 import ir, synthetic, witness_gpu
from zokrates_b200._lib import Context, ZkbError, fr_array
from zokrates_b200.ir import Constraint, Directive, LinComb, Parameter, Prog, QuadComb, Variable
from oracle.ff import BN254


def _programs():
    V = Variable
    x, y, t, u, out = V.new(0), V.new(1), V.new(2), V.new(3), V.public(0)
    yield "factorize", Prog([Parameter.private_(x), Parameter.public(y)], 0, [ir.constraint(x, x, y)]), [337, 113569]
    yield "chain", Prog([Parameter.private_(x), Parameter.public(y)], 1, [
        ir.constraint(x, y, t),                                            # t = x*y           (level 1)
        ir.constraint(LinComb([(t, 3), (V.one(), 5)]), x, u),              # u = (3t+5)*x      (level 2)
        ir.constraint(x, x, V.new(9)),                                     # independent       (level 1)
        Constraint(QuadComb(LinComb([(u, 1), (V.new(9), 2)]), LinComb.one()), LinComb.from_var(out)),   # out (level 3)
        Constraint(QuadComb(LinComb.from_var(out), LinComb.one()), LinComb([(u, 1), (V.new(9), 2)])),  # a check
    ]), [7, 11]
    yield "empty", Prog([], 0, []), []


@pytest.mark.parametrize("case", range(3))
def test_matches_interpreter_emu(case, emu_lib):
    name, prog, inputs = list(_programs())[case]
    ref = ir.Interpreter().execute(prog, inputs)
    got = witness_gpu.generate_witness(prog, inputs, lib=emu_lib)
    assert got.values == ref.values, name


def test_random_programs_and_levels_emu(emu_lib):
    for seed in range(4):
        oprog, pprog, inputs = rand_prog_pair(BN254, 40, 2, 3, seed=seed, curve_name="bn128")
        if any(isinstance(s, Directive) for s in pprog.statements):
            continue
        ref = ir.Interpreter().execute(pprog, inputs)
        assert witness_gpu.generate_witness(pprog, inputs, lib=emu_lib).values == ref.values


def test_unsatisfied_and_errors_emu(emu_lib):
    x, y = Variable.new(0), Variable.new(1)
    prog = Prog([Parameter.private_(x), Parameter.public(y)], 0, [ir.constraint(x, x, y)])
    with pytest.raises(ir.UnsatisfiedConstraint):
        witness_gpu.generate_witness(prog, [3, 10], lib=emu_lib)
    with pytest.raises(ValueError, match="WrongInputCount"):
        witness_gpu.generate_witness(prog, [3], lib=emu_lib)
    with pytest.raises(ZkbError, match="signature"):          # a malformed directive is refused when the program is loaded
        witness_gpu.generate_witness(Prog([Parameter.private_(x)], 0, [Directive([], [y], "Xor")]), [1], lib=emu_lib)


def test_programs_with_directives_route_through_the_front_door_emu(emu_lib):
    """generate_witness / prove_from_inputs on programs WITH solver directives: same witness as the interpreter mirror, same proof
    as interpreter + B200.generate_proof."""
    import io
    from tests.test_prog_native import solver_program
    from zokrates_b200 import backend, rng as prng
    prog, inputs = solver_program("bn128", 8), [201, 77, 5]
    ref = ir.Interpreter().execute(prog, inputs)
    assert witness_gpu.generate_witness(prog, inputs, lib=emu_lib).values == ref.values
    kp = backend.B200.setup(prog, [3, 5, 7, 11, 13, 17, 19], lib=emu_lib)
    want = backend.B200.generate_proof(prog, ref, io.BytesIO(kp.pk), prng.get_rng_from_entropy("d"), lib=emu_lib)
    got = witness_gpu.prove_from_inputs(prog, inputs, io.BytesIO(kp.pk), prng.get_rng_from_entropy("d"), lib=emu_lib)
    assert got.to_tagged_json() == want.to_tagged_json()


def test_inputs_to_proof_without_leaving_the_device_emu(emu_lib):
    """witness_eval leaves z resident, prove_resident consumes it: same proof JSON as interpreter + generate_proof."""
    import io
    from zokrates_b200 import backend, rng as prng
    for name, prog, inputs in _programs():
        kp = backend.B200.setup(prog, [3, 5, 7, 11, 13, 17, 19], lib=emu_lib)
        ref = backend.B200.generate_proof(prog, ir.Interpreter().execute(prog, inputs), io.BytesIO(kp.pk),
                                          prng.get_rng_from_entropy("resident"), lib=emu_lib)
        got = witness_gpu.prove_from_inputs(prog, inputs, io.BytesIO(kp.pk), prng.get_rng_from_entropy("resident"), lib=emu_lib)
        assert got.to_tagged_json() == ref.to_tagged_json(), name
        assert backend.B200.verify(kp.vk, got), name


def _synthetic_roundtrip(ctx, n):
    """synthetic circuit: forget every computed variable, let the device recompute them, compare with the generator's z;
    then the satisfaction check accepts z and names the first broken row after one value is changed."""
    r1, z = synthetic.make("bn128", n, seed=11)
    h = ctx.r1cs_load(r1.num_constraints, r1.num_instance, r1.num_witness, r1.matrices())
    m0 = r1.num_variables - n
    level_ptr, rows, out_var = witness_gpu.levelize(r1, range(m0))
    assert len(level_ptr) - 1 >= 2 and level_ptr[-1] == n and (out_var != witness_gpu.CHECK).all()
    z0 = z.copy(); z0[m0:] = 0
    got = ctx.witness_eval(h, z0, level_ptr, rows, out_var)
    assert np.array_equal(got, z)
    assert ctx.r1cs_check(h) is None                       # the assignment stayed resident
    assert ctx.r1cs_check(h, z) is None
    bad = z.copy(); bad[m0 + n // 2, 0] ^= np.uint64(1)
    first = ctx.r1cs_check(h, bad)
    assert first is not None and first <= n // 2
    with pytest.raises(ZkbError) as e:
        ctx.witness_eval(h, z0, level_ptr, rows, np.where(np.arange(n) == 5, witness_gpu.CHECK, out_var).astype(np.uint32))
    assert e.value.code == 5                               # row 5 checked against a variable nobody assigned
    ctx.r1cs_free(h)


def test_synthetic_roundtrip_emu(emu_lib):
    _synthetic_roundtrip(Context(0, 0, emu_lib), 300)


def test_wavefront_levelizer_matches_the_row_loop():
    for n, seed in ((300, 11), (3000, 3)):
        r1, z = synthetic.make("bn128", n, seed=seed)
        m0 = r1.num_variables - n
        slow = witness_gpu.levelize(r1, range(m0))
        fast = witness_gpu.levelize_wavefront(r1, range(m0))
        assert fast is not None and all(np.array_equal(a, b) for a, b in zip(slow, fast))
        assert all(np.array_equal(a, b) for a, b in zip(witness_gpu.levels_for(r1, range(m0)), slow))
    # a row that reads a variable nobody defines: the wavefront gives up, the row loop names the row
    r1, z = synthetic.make("bn128", 50, seed=1)
    assert witness_gpu.levelize_wavefront(r1, range(3)) is None
    with pytest.raises(KeyError):
        witness_gpu.levelize(r1, range(3))


def _with_empty_rows(r1, where, rs):
    """Copy of `r1` with rows `x * 0 == 0`-style (empty B and C combinations) inserted at the given positions."""
    from zokrates_b200.r1cs import R1CS
    N = r1.num_constraints
    keep = np.ones(N + len(where), dtype=bool)
    keep[np.asarray(where) + np.arange(len(where))] = False        # positions of the inserted rows in the new system
    mats = []
    for k, (rp, col, val) in enumerate(r1.matrices()):
        lens = np.zeros(N + len(where), dtype=np.uint64)
        lens[keep] = (rp[1:] - rp[:-1]).astype(np.uint64)
        col, val = col.copy(), val.copy()
        if k == 0:                                                 # the inserted rows read an input on the A side only
            extra_c = np.full(len(where), 1, dtype=np.uint32)
            extra_v = np.zeros((len(where), 4), dtype=np.uint64); extra_v[:, 0] = 1
            pos = np.cumsum(lens)[~keep].astype(np.int64) + np.arange(len(where))
            col = np.insert(col, pos - np.arange(len(where)), extra_c)
            val = np.insert(val, pos - np.arange(len(where)), extra_v, axis=0)
            lens[~keep] = 1
        nrp = np.zeros(N + len(where) + 1, dtype=np.uint64)
        np.cumsum(lens, out=nrp[1:])
        mats.append((nrp, col, val))
    return R1CS(r1.curve, N + len(where), r1.num_instance, r1.num_witness, *mats)


def test_wavefront_levelizer_with_empty_combinations():
    """rows whose B / C combinations are empty — in the middle and TRAILING — must not cut the last operand
    off the preceding row (np.reduceat on clipped start offsets did); levels equal the row loop's at N > 2048."""
    rs = np.random.RandomState(5)
    r1, z = synthetic.make("bn128", 2500, seed=7)
    m0 = r1.num_variables - 2500
    for where in ([2500], [2500, 2500, 2500], [0, 17, 1200, 2500], sorted(rs.randint(0, 2501, size=40).tolist())):
        r2 = _with_empty_rows(r1, where, rs)
        slow = witness_gpu.levelize(r2, range(m0))
        fast = witness_gpu.levelize_wavefront(r2, range(m0))
        assert fast is not None and all(np.array_equal(a, b) for a, b in zip(slow, fast)), where
    # the advisor's shape: u = x * (x + deep) as the last real row, then `x * 0 == 0` rows (empty B and C)
    from zokrates_b200.r1cs import R1CS
    r1, z = synthetic.make("bn128", 2100, seed=3)
    m = r1.num_variables
    one = np.array([[1, 0, 0, 0]], dtype=np.uint64)
    (ap, ac, av), (bp, bc, bv), (cp, cc, cv) = r1.matrices()
    deep = m - 1                                                   # the variable the last synthetic row assigned
    A = (np.concatenate([ap, [ap[-1] + 1, ap[-1] + 2, ap[-1] + 3]]).astype(np.uint64), np.concatenate([ac, [1, 1, 1]]).astype(np.uint32),
         np.concatenate([av, one, one, one]))
    B = (np.concatenate([bp, [bp[-1] + 2, bp[-1] + 2, bp[-1] + 2]]).astype(np.uint64), np.concatenate([bc, [1, deep]]).astype(np.uint32),
         np.concatenate([bv, one, one]))
    Cm = (np.concatenate([cp, [cp[-1] + 1, cp[-1] + 1, cp[-1] + 1]]).astype(np.uint64), np.concatenate([cc, [m]]).astype(np.uint32),
          np.concatenate([cv, one]))
    r2 = R1CS(r1.curve, 2103, r1.num_instance, r1.num_witness + 1, A, B, Cm)
    slow = witness_gpu.levelize(r2, range(m - 2100))
    fast = witness_gpu.levelize_wavefront(r2, range(m - 2100))
    assert fast is not None and all(np.array_equal(a, b) for a, b in zip(slow, fast))


def test_wavefront_levelizer_rejects_read_before_write():
    """a row that reads a variable assigned by a LATER row is an error in the reference interpreter
    (lookup of a missing value); the wavefront must not quietly reorder it."""
    from zokrates_b200.r1cs import R1CS
    r1, z = synthetic.make("bn128", 2200, seed=9)
    m0 = r1.num_variables - 2200
    mats = [list(m) for m in r1.matrices()]
    # swap rows 100 and 2100 in all three matrices: row 100 now reads variables defined ~2000 rows later
    perm = np.arange(2200); perm[100], perm[2100] = 2100, 100
    new = []
    for rp, col, val in mats:
        lens = (rp[1:] - rp[:-1]).astype(np.int64)
        starts = rp[:-1].astype(np.int64)
        idx = np.concatenate([np.arange(starts[r], starts[r] + lens[r]) for r in perm])
        nrp = np.zeros(2201, dtype=np.uint64); np.cumsum(lens[perm], out=nrp[1:])
        new.append((nrp, col[idx], val[idx]))
    r2 = R1CS(r1.curve, 2200, r1.num_instance, r1.num_witness, *new)
    assert witness_gpu.levelize_wavefront(r2, range(m0)) is None
    with pytest.raises(KeyError):
        witness_gpu.levels_for(r2, range(m0))


@pytest.mark.gpu
def test_synthetic_roundtrip_gpu(gpu_lib):
    ctx = Context(0, 0, gpu_lib)
    _synthetic_roundtrip(ctx, 20000)
    for _, prog, inputs in _programs():
        assert witness_gpu.generate_witness(prog, inputs, ctx=ctx).values == ir.Interpreter().execute(prog, inputs).values
    ctx.close()
