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
    with pytest.raises(NotImplementedError):
        witness_gpu.generate_witness(Prog([Parameter.private_(x)], 0, [Directive([], [y], "Xor")]), [1], lib=emu_lib)


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


@pytest.mark.gpu
def test_synthetic_roundtrip_gpu(gpu_lib):
    ctx = Context(0, 0, gpu_lib)
    _synthetic_roundtrip(ctx, 20000)
    for _, prog, inputs in _programs():
        assert witness_gpu.generate_witness(prog, inputs, ctx=ctx).values == ir.Interpreter().execute(prog, inputs).values
    ctx.close()
