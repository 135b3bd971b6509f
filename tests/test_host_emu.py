"""CPU tier: the kernel bodies and the whole engine orchestration, stepped on the host by the
emulation build (tests/host_emu/libzkb_emu.so, -DZKB_EMU), checked against the oracle.  The same
assertions run on the real GPU in tests/test_gpu_parity.py."""
import io
import random

import numpy as np
import pytest

from oracle import ark, ir as oir
from oracle.ff import BLS12_381, BN254, g1_group, g2_group
from tests.util import proof_bytes, rand_prog_pair
from zokrates_b200 import backend, proof as pproof, r1cs as pr1cs, rng as prng
from zokrates_b200._lib import Context, ZkbError, fr_array, fr_from_array

CURVES = [(0, BN254), (1, BLS12_381)]


@pytest.fixture(scope="module", params=CURVES, ids=lambda p: p[1].name)
def cc(request, emu_lib):
    cid, c = request.param
    return cid, c, Context(cid, 0, emu_lib)


def test_field_ops(cc):
    cid, c, ctx = cc
    rnd = random.Random(1)
    for field, mod, nl in ((0, c.r, 4), (1, c.p, c.fq_bytes // 8)):
        a = [0, 1, mod - 1, mod - 2] + [rnd.randrange(mod) for _ in range(60)]
        b = [mod - 1, 0, mod - 1, 2] + [rnd.randrange(mod) for _ in range(60)]
        A, B = fr_array(a, nl), fr_array(b, nl)
        assert fr_from_array(ctx.field_op(field, 0, A, B)) == [x * y % mod for x, y in zip(a, b)]
        assert fr_from_array(ctx.field_op(field, 1, A, B)) == [(x + y) % mod for x, y in zip(a, b)]
        assert fr_from_array(ctx.field_op(field, 2, A, B)) == [(x - y) % mod for x, y in zip(a, b)]
        assert fr_from_array(ctx.field_op(field, 3, A[2:12], None)) == [pow(x, -1, mod) for x in a[2:12]]


@pytest.mark.parametrize("log_n", [0, 1, 2, 3, 4, 6, 7])
def test_ntt_all_modes(cc, log_n):
    cid, c, ctx = cc
    rnd = random.Random(log_n)
    n = 1 << log_n
    x = [rnd.randrange(c.r) for _ in range(n)]
    d = ark.Domain(c, n)
    assert fr_from_array(ctx.ntt(fr_array(x))) == d.fft(x)
    assert fr_from_array(ctx.ntt(fr_array(x), inverse=True)) == d.ifft(x)
    assert fr_from_array(ctx.ntt(fr_array(x), coset=True)) == d.coset_fft(x)
    assert fr_from_array(ctx.ntt(fr_array(x), inverse=True, coset=True)) == d.coset_ifft(x)


@pytest.mark.parametrize("n", [0, 1, 2, 5, 33, 150])
def test_msm_g1_edge_cases(cc, n):
    """empty input, zero / one / r-1 scalars, infinity points, the same point twice in one bucket."""
    cid, c, ctx = cc
    rnd = random.Random(100 + n)
    G1 = g1_group(c)
    pts = [G1.mul(c.g1, rnd.randrange(1, c.r)) for _ in range(n)]
    sc = [rnd.choice([0, 1, 2, c.r - 1, rnd.randrange(c.r), rnd.randrange(1 << 20)]) for _ in range(n)]
    if n > 3:
        pts[2] = None
        pts[3] = pts[1]
        sc[3] = sc[1]
    got = ctx.msm(1, b"".join(ark.ser_g1(c, p) for p in pts), fr_array(sc))
    assert got == ark.ser_g1(c, G1.msm_naive(pts, sc))


def test_msm_skewed_big_bucket(cc):
    """90 % unit scalars over 8 distinct points: one huge bucket, cut by many chunk borders (level >= 2)."""
    cid, c, ctx = cc
    rnd = random.Random(7)
    G1 = g1_group(c)
    base = [G1.mul(c.g1, rnd.randrange(1, c.r)) for _ in range(8)]
    n = 700
    sc = [1 if rnd.random() < 0.9 else rnd.randrange(c.r) for _ in range(n)]
    got = ctx.msm(1, b"".join(ark.ser_g1(c, base[i % 8]) for i in range(n)), fr_array(sc))
    sums = [sum(sc[j::8]) % c.r for j in range(8)]
    assert got == ark.ser_g1(c, G1.msm_naive(base, sums))


@pytest.mark.parametrize("n", [0, 1, 3, 24])
def test_msm_g2(cc, n):
    cid, c, ctx = cc
    rnd = random.Random(200 + n)
    G2 = g2_group(c)
    pts = [G2.mul(c.g2, rnd.randrange(1, c.r)) for _ in range(n)]
    sc = [rnd.choice([0, 1, rnd.randrange(c.r)]) for _ in range(n)]
    if n > 2:
        pts[1] = None
    assert ctx.msm(2, b"".join(ark.ser_g2(c, p) for p in pts), fr_array(sc)) == ark.ser_g2(c, G2.msm_naive(pts, sc))


@pytest.mark.parametrize("shape", [(1, 1, 1), (5, 2, 1), (13, 0, 3), (30, 1, 2), (100, 1, 2)], ids=str)
def test_prove_through_backend_mirror(cc, shape, emu_lib):
    """`B200::generate_proof` (host mirror -> C ABI) equals the oracle's prover byte for byte, the
    trapdoor prediction, and verifies; setup output equals the oracle's proving key bytes."""
    cid, c, ctx = cc
    ncons, npub, npriv = shape
    if cid == 1 and shape in ((5, 2, 1), (30, 1, 2)):
        pytest.skip("BLS12-381 runs three of the five shapes (the python oracle's 381-bit pairing dominates the CPU tier)")
    oprog, pprog, inputs = rand_prog_pair(c, ncons, npub, npriv, seed=hash(shape) & 0xFFFF, curve_name=c.name)
    ow = oir.execute(c, oprog, inputs)
    from zokrates_b200.ir import Interpreter
    pw = Interpreter().execute(pprog, inputs)
    assert {v.id: x for v, x in pw.values.items()} == ow
    r1cs_o, z = ark.synthesize(oprog, ow)
    r1cs_p = pr1cs.synthesize(pprog)
    assert (r1cs_p.num_constraints, r1cs_p.num_instance, r1cs_p.num_witness) == (r1cs_o.num_constraints, r1cs_o.num_instance, r1cs_o.num_witness)
    assert fr_from_array(r1cs_p.assignment(pw)) == z
    rnd = random.Random(5)
    tdv = [rnd.randrange(1, c.r) for _ in range(7)]
    kp = backend.B200.setup(pprog, tdv, lib=emu_lib)
    td = ark.Trapdoor(*tdv)
    if ncons <= 13:
        assert kp.pk == ark.pk_serialize(c, ark.setup(c, r1cs_o, td))
    proof = backend.B200.generate_proof(pprog, pw, io.BytesIO(kp.pk), prng.get_rng_from_entropy("seed"), lib=emu_lib)
    orng = ark.rng_from_entropy("seed")
    r, s = ark.fr_rand(c, orng), ark.fr_rand(c, orng)
    exp = ark.trapdoor_expected_proof(c, r1cs_o, td, z, r, s)
    assert proof.to_raw() == proof_bytes(c, exp)
    (oproof, oinputs) = ark.generate_proof(c, oprog, ow, kp.pk, ark.rng_from_entropy("seed"))
    assert proof.to_tagged_json() == ark.tagged_proof_json(c, oproof, oinputs)
    if ncons <= 5:
        assert ark.verify(c, ark.pk_deserialize(c, kp.pk), oinputs, oproof)
        assert backend.B200.verify(kp.vk, proof)        # the product's own host verifier (Backend::verify)


def test_sharded_partials_equal_single(cc):
    cid, c, ctx = cc
    oprog, pprog, inputs = rand_prog_pair(c, 21, 1, 2, seed=77, curve_name=c.name)
    ow = oir.execute(c, oprog, inputs)
    r1cs_o, z = ark.synthesize(oprog, ow)
    r1cs_p = pr1cs.synthesize(pprog)
    h = ctx.r1cs_load(r1cs_p.num_constraints, r1cs_p.num_instance, r1cs_p.num_witness, r1cs_p.matrices())
    pk = ctx.setup(h, [3, 5, 7, 11, 13, 17, 19])
    Z = fr_array(z)
    single = ctx.prove(ctx.pk_load(pk), h, Z, 111, 222)
    for world in (2, 3, 5):
        parts = [ctx.prove_partial(ctx.pk_load(pk, rank, world), h, Z) for rank in range(world)]
        assert ctx.finalize(ctx.pk_load(pk), np.concatenate(parts), world, 111, 222) == single
    # resident assignment path
    ctx.set_assignment(h, Z)
    assert ctx.prove_resident(ctx.pk_load(pk), h, 111, 222) == single


def test_error_paths(cc):
    cid, c, ctx = cc
    with pytest.raises(ZkbError) as e:
        ctx.pk_load(b"\x00" * 100)
    assert e.value.code == 2          # ZKB_E_FORMAT: truncated key
    with pytest.raises(ZkbError) as e:
        ctx.prove(12345, 67890, np.zeros((3, 4), dtype=np.uint64), 1, 2)
    assert e.value.code == 1          # ZKB_E_ARG: unknown handles
    oprog, pprog, inputs = rand_prog_pair(c, 3, 1, 1, seed=1, curve_name=c.name)
    r1 = pr1cs.synthesize(pprog)
    h = ctx.r1cs_load(r1.num_constraints, r1.num_instance, r1.num_witness, r1.matrices())
    pk = ctx.setup(h, [3, 5, 7, 11, 13, 17, 19])
    with pytest.raises(ZkbError):     # trailing garbage after the key
        ctx.pk_load(pk + b"\x00")
    o2, p2, _ = rand_prog_pair(c, 9, 1, 1, seed=2, curve_name=c.name)
    r2 = pr1cs.synthesize(p2)
    h2 = ctx.r1cs_load(r2.num_constraints, r2.num_instance, r2.num_witness, r2.matrices())
    with pytest.raises(ZkbError):     # key of another circuit
        ctx.prove(ctx.pk_load(pk), h2, np.zeros((r2.num_variables, 4), dtype=np.uint64), 1, 2)


@pytest.mark.parametrize("case", range(7))
def test_reference_edge_programs(case, emu_lib):
    """empty / identity / public identity / no arguments / `+ one` / unordered variables / public output
    (the program shapes of the reference's own backend tests)."""
    from tests.util import check_backend_roundtrip, reference_edge_programs
    check_backend_roundtrip(emu_lib, *reference_edge_programs()[case])


@pytest.mark.parametrize("dist", ["uniform", "bits"])
def test_prove_synthetic_vs_c_oracle(dist, emu_lib, oracle_c):
    """300-constraint synthetic circuit (both witness distributions: full-width and 90 % bits, the latter taking the
    sparse 16-window path) against the ark-equivalent C prover; setup bytes equal the C oracle's key."""
    from zokrates_b200 import synthetic
    ctx = Context(0, 0, emu_lib)
    r1, z = synthetic.make("bn128", 300, distribution=dist)
    h = ctx.r1cs_load(r1.num_constraints, r1.num_instance, r1.num_witness, r1.matrices())
    td = [3, 5, 7, 11, 13, 17, 19]
    pk = ctx.setup(h, td)
    assert pk == oracle_c.setup(0, r1, td)
    ref, _ = oracle_c.prove(0, pk, r1, z, 111, 222, 32)
    assert ctx.prove(ctx.pk_load(pk), h, z, 111, 222) == ref
    parts = [ctx.prove_partial(ctx.pk_load(pk, k, 2), h, z) for k in range(2)]
    assert ctx.finalize(ctx.pk_load(pk), np.concatenate(parts), 2, 111, 222) == ref


@pytest.mark.parametrize("log_n,max_s", [(10, 10), (11, 10), (12, 10), (13, 10), (10, 5), (12, 6), (13, 5), (14, 7), (15, 5)])
def test_ntt_tiled_passes(cc, log_n, max_s):
    """Shared-memory tile passes (ntt_block_body): one pass (2^10), two balanced passes, and — with the per-pass stage
    limit lowered — three passes with interleaved groups, against the python oracle in all four modes; the register-pass
    schedule (tile path switched off) must give the same vectors."""
    cid, c, ctx = cc
    from zokrates_b200._lib import OPT_NTT_MAX_S, OPT_NTT_TILE_MIN
    ctx.set_option(OPT_NTT_MAX_S, max_s)
    ctx.set_option(OPT_NTT_TILE_MIN, 10)
    rnd = random.Random(1000 + log_n)
    n = 1 << log_n
    x = [rnd.randrange(c.r) for _ in range(n)]
    d = ark.Domain(c, n)
    X = fr_array(x)
    got = [ctx.ntt(X), ctx.ntt(X, inverse=True), ctx.ntt(X, coset=True), ctx.ntt(X, inverse=True, coset=True)]
    assert fr_from_array(got[0]) == d.fft(x)
    assert fr_from_array(got[1]) == d.ifft(x)
    assert fr_from_array(got[2]) == d.coset_fft(x)
    assert fr_from_array(got[3]) == d.coset_ifft(x)
    ctx.set_option(OPT_NTT_TILE_MIN, 30)
    legacy = [ctx.ntt(X), ctx.ntt(X, inverse=True), ctx.ntt(X, coset=True), ctx.ntt(X, inverse=True, coset=True)]
    ctx.set_option(OPT_NTT_TILE_MIN, 10)
    ctx.set_option(OPT_NTT_MAX_S, 10)
    for a, b in zip(got, legacy):
        assert np.array_equal(a, b)


def test_pipelined_submit_collect(cc):
    """zkb_groth16_prove_submit / _collect: two proofs in flight (different blinding scalars) give the bytes of
    the one-at-a-time calls, in either collection order; a third submit is refused; keys and matrices of a proof in flight
    cannot be freed; partial tickets combine through zkb_groth16_finalize."""
    from zokrates_b200 import synthetic
    from zokrates_b200._lib import ZkbError
    cid, c, ctx = cc
    r1, z1 = synthetic.make(c.name, 70, seed=5)
    z2 = z1.copy()
    h = ctx.r1cs_load(r1.num_constraints, r1.num_instance, r1.num_witness, r1.matrices())
    pk = ctx.setup(h, [3, 5, 7, 11, 13, 17, 19])
    pkh = ctx.pk_load(pk)
    ref1 = ctx.prove(pkh, h, z1, 111, 222)
    ref2 = ctx.prove(pkh, h, z2, 333, 444)
    assert ref1 != ref2
    t1 = ctx.prove_submit(pkh, h, z1, 111, 222)
    t2 = ctx.prove_submit(pkh, h, z2, 333, 444)
    with pytest.raises(ZkbError):
        ctx.prove_submit(pkh, h, z1, 1, 2)
    with pytest.raises(ZkbError):
        ctx.pk_free(pkh)
    with pytest.raises(ZkbError):
        ctx.r1cs_free(h)
    assert ctx.prove_collect(t2) == ref2
    assert ctx.prove_collect(t1) == ref1
    with pytest.raises(ZkbError):
        ctx.prove_collect(t1)
    # resident assignment + partial tickets
    ctx.set_assignment(h, z1)
    ta = ctx.prove_submit(pkh, h, None)
    tb = ctx.prove_submit(pkh, h, z2)
    pa, pb = ctx.prove_collect_partial(ta), ctx.prove_collect_partial(tb)
    assert ctx.finalize(pkh, pa, 1, 111, 222) == ref1
    assert ctx.finalize(pkh, pb, 1, 333, 444) == ref2
    tp = ctx.prove_submit(pkh, h, z1)
    with pytest.raises(ZkbError):
        ctx.prove_collect(tp)                                    # a partial ticket has no r, s; it stays collectable
    assert ctx.finalize(pkh, ctx.prove_collect_partial(tp), 1, 111, 222) == ref1
    # the legacy begin / end pair and the async pair interleave with a submitted proof
    tk = ctx.prove_submit(pkh, h, z2, 333, 444)
    t3, ptrs, nbytes = ctx.prove_begin_async(pkh, h, z1, 7)
    ctx.prove_end_async(t3)
    assert ctx.prove_collect(tk) == ref2
    assert ctx.finalize(pkh, ctx.prove_collect_partial(t3), 1, 111, 222) == ref1
    ctx.pk_free(pkh)
    ctx.r1cs_free(h)
