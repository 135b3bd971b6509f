"""Batch-affine rounds in front of the bucket accumulation (csrc/msm_affine.cuh): same sums as the direct XYZZ path and as the
oracle, including the cases the shared inversion must survive — equal points in one bucket (doubling), P and -P (cancellation),
points at infinity, one bucket holding everything.  CPU tier: the block passes (forward products, product tree, one inversion,
backward pass) run in the host emulation, thread after thread; -m gpu: the device kernel at sizes with several blocks."""
import random

import numpy as np
import pytest

from oracle import ark
from oracle.ff import BLS12_381, BN254, g1_group, g2_group
from zokrates_b200._lib import (OPT_BATCH_AFFINE, OPT_BATCH_AFFINE_MIN_LOG, Context, fr_array)

CURVES = {0: BN254, 1: BLS12_381}


def _points(c, group, n, rnd):
    G = g1_group(c) if group == 1 else g2_group(c)
    gen = c.g1 if group == 1 else c.g2
    base = [G.mul(gen, rnd.randrange(1, c.r)) for _ in range(min(n, 24))]
    return G, [base[i % len(base)] if i >= len(base) and rnd.random() < 0.5 else G.mul(gen, rnd.randrange(1, 1 << 40)) if i >= len(base) else base[i]
               for i in range(n)]


def _ser(c, group, pts):
    ser = ark.ser_g1 if group == 1 else ark.ser_g2
    return b"".join(ser(c, p) for p in pts)


def _cases(c, group, rnd):
    G, pts = _points(c, group, 90, rnd)
    q = c.r
    uni = [rnd.randrange(q) for _ in pts]
    yield "uniform", pts, uni
    yield "bits", pts, [rnd.choice([0, 1, 1, 2, rnd.randrange(q)]) for _ in pts]
    same = rnd.randrange(q)
    yield "one-bucket", pts, [same] * len(pts)                                  # every window: one bucket holds all points
    dup = pts[:10] * 6 + [None, None] + pts[10:20]                                # equal points with equal scalars: doublings; infinity points
    sc = ([7] * 10 + [9] * 10) * 3 + [5, 6] + [rnd.randrange(q) for _ in range(10)]
    yield "duplicates+infinity", dup, sc
    neg = []
    for p in pts[:16]:
        neg += [p, G.neg(p)]
    yield "cancellation", neg + pts[16:30], [12345] * 32 + [rnd.randrange(q) for _ in range(14)]   # P + (-P) inside a bucket


def _run(lib, cid, group, rounds_list=(1, 3, 6)):
    c = CURVES[cid]
    rnd = random.Random(100 * cid + group)
    ctx = Context(cid, 0, lib)
    ctx.set_option(OPT_BATCH_AFFINE_MIN_LOG, 0)
    G = g1_group(c) if group == 1 else g2_group(c)
    for name, pts, sc in _cases(c, group, rnd):
        data, scal = _ser(c, group, pts), fr_array(sc)
        want = _ser(c, group, [G.msm_naive(pts, sc)])
        ctx.set_option(OPT_BATCH_AFFINE, 0)
        assert ctx.msm(group, data, scal) == want, (name, "direct")
        for rounds in rounds_list:
            ctx.set_option(OPT_BATCH_AFFINE, rounds)
            assert ctx.msm(group, data, scal) == want, (name, rounds)


@pytest.mark.parametrize("cid,group", [(0, 1), (0, 2), (1, 1), (1, 2)])
def test_batch_affine_msm_emu(cid, group, emu_lib):
    _run(emu_lib, cid, group)


def test_batch_affine_proof_emu(emu_lib):
    """A whole proof (views, window tables off/on, five MSMs) is unchanged by the affine rounds."""
    from tests.util import rand_prog_pair
    from zokrates_b200 import ir as pir
    from zokrates_b200.r1cs import synthesize
    from zokrates_b200._lib import OPT_TABLE_MIN_LOG
    c = BN254
    oprog, pprog, inputs = rand_prog_pair(c, 120, 2, 3, seed=7, curve_name="bn128")
    w = pir.Interpreter().execute(pprog, inputs)
    r1 = synthesize(pprog)
    ctx = Context(0, 0, emu_lib)
    h = ctx.r1cs_load(r1.num_constraints, r1.num_instance, r1.num_witness, r1.matrices())
    pk = ctx.setup(h, [3, 5, 7, 11, 13, 17, 19])
    z = r1.assignment(w)
    proofs = set()
    for tables_min, rounds in ((14, 0), (14, 3), (0, 0), (0, 2)):
        ctx.set_option(OPT_TABLE_MIN_LOG, tables_min)
        ctx.set_option(OPT_BATCH_AFFINE_MIN_LOG, 0)
        ctx.set_option(OPT_BATCH_AFFINE, rounds)
        pkh = ctx.pk_load(pk)
        proofs.add(ctx.prove(pkh, h, z, 1234567, 7654321))
        ctx.pk_free(pkh)
    assert len(proofs) == 1


@pytest.mark.gpu
@pytest.mark.parametrize("cid,group", [(0, 1), (0, 2), (1, 1), (1, 2)])
def test_batch_affine_msm_gpu(cid, group, gpu_lib):
    _run(gpu_lib, cid, group, rounds_list=(1, 3))


@pytest.mark.gpu
@pytest.mark.parametrize("cid", [0, 1])
def test_batch_affine_large_gpu(cid, gpu_lib, oracle_c):
    """2^15 distinct points (GPU-made key), uniform and 0/1-heavy scalars: many blocks per round, against the C oracle."""
    from zokrates_b200 import synthetic
    c = CURVES[cid]
    ctx = Context(cid, 0, gpu_lib)
    name = "bn128" if cid == 0 else "bls12_381"
    lg = 15
    r1, z = synthetic.make_layered(ctx, name, (1 << lg) - 2)
    h = ctx.r1cs_load(r1.num_constraints, r1.num_instance, r1.num_witness, r1.matrices())
    pk = ctx.setup(h, [11, 22, 33, 44, 5555, 3, 7])
    ctx.r1cs_free(h)
    g1, g2 = 2 * c.fq_bytes, 4 * c.fq_bytes
    m = r1.num_variables
    off = g1 + 3 * g2 + 8 + 2 * g1 + 2 * g1 + (8 + m * g1) * 2 + 8 + m * g2 + 8          # start of h_query (ni = 2)
    n = (1 << lg) - 1
    pts = bytes(pk[off:off + n * g1])
    rs = np.random.RandomState(5)
    for dist in ("uniform", "bits"):
        sc = rs.randint(0, 1 << 62, size=(n, 4)).astype(np.uint64)
        sc[:, 3] &= np.uint64((1 << 60) - 1)
        if dist == "bits":
            small = rs.rand(n) < 0.9
            sc[small, 1:] = 0
            sc[small, 0] = rs.randint(0, 2, size=int(small.sum())).astype(np.uint64)
        want = oracle_c.msm(cid, 1, pts, sc, c.fq_bytes)
        for rounds in (0, 3):
            ctx.set_option(OPT_BATCH_AFFINE, rounds)
            ctx.set_option(OPT_BATCH_AFFINE_MIN_LOG, 0)
            assert ctx.msm(1, pts, sc) == want, (dist, rounds)
