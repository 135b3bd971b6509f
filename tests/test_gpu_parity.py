"""GPU tier (-m gpu): parity of the sm_100a kernels against the oracle THROUGH the C ABI of libzkb200.so.

Small sizes compare with the python big-int oracle, mid sizes with the C restatement (both bit-exact:
field elements, affine points and proof bytes are canonical), full BASELINE sizes through
size-independent properties (NTT round trip, MSM linearity, proof equality across shardings).
Nothing here reads /root/reference."""
import io
import random

import numpy as np
import pytest

from oracle import ark, ir as oir
from oracle.ff import BLS12_381, BN254, g1_group, g2_group
from tests.util import proof_bytes, rand_prog_pair
from zokrates_b200 import backend, r1cs as pr1cs, rng as prng, synthetic
from zokrates_b200._lib import Context, ZkbError, fr_array, fr_from_array

pytestmark = pytest.mark.gpu
CURVES = [(0, BN254), (1, BLS12_381)]


@pytest.fixture(scope="module", params=CURVES, ids=lambda p: p[1].name)
def cc(request, gpu_lib):
    cid, c = request.param
    ctx = Context(cid, 0, gpu_lib)
    yield cid, c, ctx
    ctx.close()


def test_field_ops_vs_bigint(cc):
    cid, c, ctx = cc
    rnd = random.Random(1)
    for field, mod, nl in ((0, c.r, 4), (1, c.p, c.fq_bytes // 8)):
        a = [0, 1, mod - 1, mod - 2] + [rnd.randrange(mod) for _ in range(4000)]
        b = [mod - 1, 0, mod - 1, 2] + [rnd.randrange(mod) for _ in range(4000)]
        A, B = fr_array(a, nl), fr_array(b, nl)
        assert fr_from_array(ctx.field_op(field, 0, A, B)) == [x * y % mod for x, y in zip(a, b)]
        assert fr_from_array(ctx.field_op(field, 1, A, B)) == [(x + y) % mod for x, y in zip(a, b)]
        assert fr_from_array(ctx.field_op(field, 2, A, B)) == [(x - y) % mod for x, y in zip(a, b)]
        assert fr_from_array(ctx.field_op(field, 3, A[2:200], None)) == [pow(x, -1, mod) for x in a[2:200]]


@pytest.mark.parametrize("log_n", [0, 1, 2, 3, 5, 8, 10])
def test_ntt_vs_bigint(cc, log_n):
    cid, c, ctx = cc
    rnd = random.Random(log_n)
    x = [rnd.randrange(c.r) for _ in range(1 << log_n)]
    d = ark.Domain(c, 1 << log_n)
    assert fr_from_array(ctx.ntt(fr_array(x))) == d.fft(x)
    assert fr_from_array(ctx.ntt(fr_array(x), inverse=True)) == d.ifft(x)
    assert fr_from_array(ctx.ntt(fr_array(x), coset=True)) == d.coset_fft(x)
    assert fr_from_array(ctx.ntt(fr_array(x), inverse=True, coset=True)) == d.coset_ifft(x)


@pytest.mark.parametrize("log_n", [13, 16, 19])
def test_ntt_vs_c_oracle(cc, oracle_c, log_n):
    cid, c, ctx = cc
    rs = np.random.RandomState(log_n)
    x = rs.randint(0, 1 << 62, size=(1 << log_n, 4)).astype(np.uint64)
    x[:, 3] &= np.uint64((1 << 60) - 1)
    for inv, coset in ((False, False), (True, True)):
        assert np.array_equal(ctx.ntt(x, inverse=inv, coset=coset), oracle_c.ntt(cid, x, inv, coset))


def test_ntt_roundtrip_full_size(cc):
    """BASELINE config 5 size (2^22 here to bound host memory): ifft(fft(x)) == x and coset variant."""
    cid, c, ctx = cc
    rs = np.random.RandomState(5)
    x = rs.randint(0, 1 << 62, size=(1 << 22, 4)).astype(np.uint64)
    x[:, 3] &= np.uint64((1 << 60) - 1)
    assert np.array_equal(ctx.ntt(ctx.ntt(x), inverse=True), x)
    assert np.array_equal(ctx.ntt(ctx.ntt(x, coset=True), inverse=True, coset=True), x)


@pytest.mark.parametrize("n", [0, 1, 2, 5, 33, 500])
def test_msm_g1_edge_cases(cc, n):
    cid, c, ctx = cc
    rnd = random.Random(100 + n)
    G1 = g1_group(c)
    pts = [G1.mul(c.g1, rnd.randrange(1, c.r)) for _ in range(min(n, 40))]
    pts = [pts[i % len(pts)] for i in range(n)] if n else []
    sc = [rnd.choice([0, 1, 2, c.r - 1, rnd.randrange(c.r), rnd.randrange(1 << 20)]) for _ in range(n)]
    if n > 3:
        pts[2] = None
        pts[3] = pts[1]
        sc[3] = sc[1]
    got = ctx.msm(1, b"".join(ark.ser_g1(c, p) for p in pts), fr_array(sc))
    assert got == ark.ser_g1(c, G1.msm_naive(pts, sc))


@pytest.mark.parametrize("n", [0, 1, 3, 60])
def test_msm_g2_vs_bigint(cc, n):
    cid, c, ctx = cc
    rnd = random.Random(200 + n)
    G2 = g2_group(c)
    base = [G2.mul(c.g2, rnd.randrange(1, c.r)) for _ in range(min(n, 12))]
    pts = [base[i % len(base)] for i in range(n)] if n else []
    sc = [rnd.choice([0, 1, rnd.randrange(c.r)]) for _ in range(n)]
    if n > 2:
        pts[1] = None
    assert ctx.msm(2, b"".join(ark.ser_g2(c, p) for p in pts), fr_array(sc)) == ark.ser_g2(c, G2.msm_naive(pts, sc))


def _key_points(ctx, c, n, seed):
    """n distinct G1/G2 points with ark encoding, taken from a GPU-made proving key (a_query / b_g2_query)."""
    r1, z = synthetic.make(c.name, n, seed=seed)
    h = ctx.r1cs_load(r1.num_constraints, r1.num_instance, r1.num_witness, r1.matrices())
    pk = ctx.setup(h, [3, 5, 7, 11, 13 + seed, 17, 19])
    ctx.r1cs_free(h)
    key = ark_pk_slices(c, pk)
    return key


def ark_pk_slices(c, pk):
    n = c.fq_bytes
    off = 2 * n + 3 * 4 * n
    cnt = int.from_bytes(pk[off:off + 8], "little"); off += 8 + cnt * 2 * n + 2 * 2 * n
    m = int.from_bytes(pk[off:off + 8], "little"); off += 8
    a = pk[off:off + m * 2 * n]; off += m * 2 * n + 8
    off += m * 2 * n + 8
    b2 = pk[off:off + m * 4 * n]
    return a, b2, m


@pytest.mark.parametrize("dist", ["uniform", "bits", "bytes"])
def test_msm_vs_c_oracle_mid(cc, oracle_c, dist):
    """2^13 pairs, three scalar distributions (uniform / 90 % {0,1} / small values): ark-equivalent CPU MSM
    must give the same affine bytes (G1 and G2)."""
    cid, c, ctx = cc
    a, b2, m = _key_points(ctx, c, (1 << 13) - 8, 1)
    rs = np.random.RandomState(3)
    sc = rs.randint(0, 1 << 62, size=(m, 4)).astype(np.uint64)
    sc[:, 3] &= np.uint64((1 << 60) - 1)
    if dist == "bits":
        small = rs.rand(m) < 0.9
        sc[small] = 0
        sc[small, 0] = rs.randint(0, 2, size=int(small.sum())).astype(np.uint64)
    elif dist == "bytes":
        sc[:] = 0
        sc[:, 0] = rs.randint(0, 256, size=m).astype(np.uint64)
    assert ctx.msm(1, a, sc) == oracle_c.msm(cid, 1, a, sc, c.fq_bytes)
    k = m // 4
    assert ctx.msm(2, b2[:k * 4 * c.fq_bytes], sc[:k]) == oracle_c.msm(cid, 2, b2[:k * 4 * c.fq_bytes], sc[:k], c.fq_bytes)


def test_msm_linearity_large(cc):
    """2^18 pairs: MSM(P, s + t) == MSM(P, s) + MSM(P, t) checked by a 2-point oracle-side addition, and
    MSM over a permutation of the pairs is unchanged."""
    cid, c, ctx = cc
    a, _, m = _key_points(ctx, c, (1 << 18) - 8, 2)
    rs = np.random.RandomState(4)
    s = rs.randint(0, 1 << 62, size=(m, 4)).astype(np.uint64); s[:, 3] &= np.uint64((1 << 59) - 1)
    t = rs.randint(0, 1 << 62, size=(m, 4)).astype(np.uint64); t[:, 3] &= np.uint64((1 << 59) - 1)
    st = fr_array([(x + y) for x, y in zip(fr_from_array(s[:2000]), fr_from_array(t[:2000]))])
    # linearity on a 2000-pair prefix (host big-int addition of scalars), permutation on the full size
    n2 = 2000 * 2 * c.fq_bytes
    G1 = g1_group(c)
    ps = ark.de_g1(c, ctx.msm(1, a[:n2], s[:2000]), 0)[0]
    pt = ark.de_g1(c, ctx.msm(1, a[:n2], t[:2000]), 0)[0]
    assert ctx.msm(1, a[:n2], st) == ark.ser_g1(c, G1.add(ps, pt))
    perm = rs.permutation(m)
    pts = np.frombuffer(a, dtype=np.uint8).reshape(m, 2 * c.fq_bytes)
    assert ctx.msm(1, pts[perm].tobytes(), s[perm]) == ctx.msm(1, a, s)


@pytest.mark.parametrize("shape", [(1, 1, 1), (5, 2, 1), (13, 0, 3), (100, 1, 2)], ids=str)
def test_generate_proof_vs_oracle(cc, shape, gpu_lib):
    """Backend mirror -> C ABI -> GPU: proof JSON equals the python oracle's prover and the trapdoor
    prediction; the proof verifies (pairing check).  Setup bytes equal the oracle's key."""
    cid, c, ctx = cc
    ncons, npub, npriv = shape
    oprog, pprog, inputs = rand_prog_pair(c, ncons, npub, npriv, seed=hash(shape) & 0xFFFF, curve_name=c.name)
    ow = oir.execute(c, oprog, inputs)
    from zokrates_b200.ir import Interpreter
    pw = Interpreter().execute(pprog, inputs)
    r1cs_o, z = ark.synthesize(oprog, ow)
    rnd = random.Random(5)
    tdv = [rnd.randrange(1, c.r) for _ in range(7)]
    kp = backend.B200.setup(pprog, tdv)
    td = ark.Trapdoor(*tdv)
    if ncons <= 13:
        assert kp.pk == ark.pk_serialize(c, ark.setup(c, r1cs_o, td))
    proof = backend.B200.generate_proof(pprog, pw, io.BytesIO(kp.pk), prng.get_rng_from_entropy("seed"))
    orng = ark.rng_from_entropy("seed")
    r, s = ark.fr_rand(c, orng), ark.fr_rand(c, orng)
    exp = ark.trapdoor_expected_proof(c, r1cs_o, td, z, r, s)
    assert proof.to_raw() == proof_bytes(c, exp)
    assert proof.to_tagged_json() == ark.tagged_proof_json(c, exp, oprog.public_inputs_values(ow))
    if ncons <= 13:
        assert ark.verify(c, ark.pk_deserialize(c, kp.pk), oprog.public_inputs_values(ow), exp)
        assert backend.B200.verify(kp.vk, proof)        # the product's own host verifier (Backend::verify)


@pytest.mark.parametrize("dist", ["uniform", "bits"])
def test_prove_vs_c_oracle_mid(cc, oracle_c, dist):
    """2^12-constraint synthetic circuit: GPU setup == C-oracle setup (bytes), witness_map equal, GPU proof
    == ark-equivalent CPU proof (bytes)."""
    cid, c, ctx = cc
    r1, z = synthetic.make(c.name, (1 << 12) - 2, distribution=dist)
    h = ctx.r1cs_load(r1.num_constraints, r1.num_instance, r1.num_witness, r1.matrices())
    td = [3, 5, 7, 11, 1234567, 17, 19]
    pk = ctx.setup(h, td)
    assert pk == oracle_c.setup(cid, r1, td)
    assert np.array_equal(ctx.witness_map(h, z, r1.domain_size), oracle_c.witness_map(cid, r1, z))
    pkh = ctx.pk_load(pk)
    proof = ctx.prove(pkh, h, z, 111, 222)
    ref, _ = oracle_c.prove(cid, pk, r1, z, 111, 222, c.fq_bytes)
    assert proof == ref
    parts = [ctx.prove_partial(ctx.pk_load(pk, rank, 4), h, z) for rank in range(4)]
    assert ctx.finalize(pkh, np.concatenate(parts), 4, 111, 222) == proof
    ctx.set_assignment(h, z)
    assert ctx.prove_resident(pkh, h, 111, 222) == proof


def test_error_paths(cc):
    cid, c, ctx = cc
    with pytest.raises(ZkbError) as e:
        ctx.pk_load(b"\x00" * 100)
    assert e.value.code == 2
    with pytest.raises(ZkbError) as e:
        ctx.prove(12345, 67890, np.zeros((3, 4), dtype=np.uint64), 1, 2)
    assert e.value.code == 1


@pytest.mark.parametrize("case", range(7))
def test_reference_edge_programs(case, gpu_lib):
    from tests.util import check_backend_roundtrip, reference_edge_programs
    check_backend_roundtrip(None, *reference_edge_programs()[case])


def test_file_level_generate_proof_tool(tmp_path, gpu_lib):
    """`zkb-generate-proof -i out -w witness -p proving.key -j proof.json -e entropy` (the file-level drop-in for
    zokrates_cli/src/ops/generate_proof.rs:95-202): program file, binary witness and ark-format key in, TaggedProof JSON
    out — equal to the python oracle's proof for the same entropy (factorize.zok of BASELINE config 1)."""
    import json
    from tools import zkb_generate_proof as tool
    from zokrates_b200 import ir, zir
    c = BN254
    a_, b_ = ir.Variable.new(0), ir.Variable.new(1)
    prog = ir.Prog([ir.Parameter.private_(a_), ir.Parameter.public(b_)], 0, [ir.constraint(a_, a_, b_)], "bn128")
    witness = ir.Interpreter().execute(prog, [337, 113569])
    td = [11, 22, 33, 44, 55555, 3, 7]
    kp = backend.B200.setup(prog, td)
    (tmp_path / "out").write_bytes(zir.write_prog(prog))
    (tmp_path / "witness").write_bytes(witness.write())
    (tmp_path / "proving.key").write_bytes(kp.pk)
    rc = tool.main(["-i", str(tmp_path / "out"), "-w", str(tmp_path / "witness"), "-p", str(tmp_path / "proving.key"),
                    "-j", str(tmp_path / "proof.json"), "-e", "file-level"])
    assert rc == 0
    text = (tmp_path / "proof.json").read_text()
    oprog = oir.Prog([(a_.id, True), (b_.id, False)], 0, [oir.Constraint([(a_.id, 1)], [(a_.id, 1)], [(b_.id, 1)])])
    ow = oir.execute(c, oprog, [337, 113569])
    r1cs_o, z = ark.synthesize(oprog, ow)
    orng = ark.rng_from_entropy("file-level")
    r, s = ark.fr_rand(c, orng), ark.fr_rand(c, orng)
    exp = ark.trapdoor_expected_proof(c, r1cs_o, ark.Trapdoor(*td), z, r, s)
    assert text == ark.tagged_proof_json(c, exp, [113569])
    assert json.loads(text)["scheme"] == "g16" and json.loads(text)["curve"] == "bn128"
    with pytest.raises(SystemExit, match="Could not open"):
        tool.main(["-i", str(tmp_path / "missing")])


def test_prove_begin_end_chain_exchange(gpu_lib):
    """zkb_groth16_prove_begin / _end on one GPU: the chain buffers are device memory a torch tensor can share
    (__cuda_array_interface__, what the NCCL broadcast uses); chains this rank skips must be supplied by the caller."""
    import torch
    from zokrates_b200.distributed import chain_tensor
    ctx = Context(0, 0, gpu_lib)
    r1, z = synthetic.make("bn128", (1 << 13) - 2)
    h = ctx.r1cs_load(r1.num_constraints, r1.num_instance, r1.num_witness, r1.matrices())
    pkh = ctx.pk_load(ctx.setup(h, [3, 5, 7, 11, 1234567, 17, 19]))
    # partial blobs hold projective points (their representation depends on the order the sort's atomics happened to
    # produce), so results are compared as finished proofs — affine, canonical
    fin = lambda partial: ctx.finalize(pkh, partial, 1, 111, 222)
    ref = fin(ctx.prove_partial(pkh, h, z))
    assert ref == ctx.prove(pkh, h, z, 111, 222)
    ptrs, nbytes = ctx.prove_begin(pkh, h, z, 7)
    assert nbytes == 32 << 13 and all(ptrs)
    assert fin(ctx.prove_end(pkh, h)) == ref
    dev = torch.device("cuda", 0)
    tb, tc = chain_tensor(ptrs[1], nbytes, dev), chain_tensor(ptrs[2], nbytes, dev)
    keep_b, keep_c = tb.clone(), tc.clone()          # coset evaluations of B z and C z (the finish step only consumes a)
    tb.zero_(); tc.zero_(); torch.cuda.synchronize()
    ctx.prove_begin(pkh, h, z, 1)                    # this "rank" computes chain a only
    assert fin(ctx.prove_end(pkh, h)) != ref
    ctx.prove_begin(pkh, h, z, 1)
    tb.copy_(keep_b); tc.copy_(keep_c); torch.cuda.synchronize()
    assert fin(ctx.prove_end(pkh, h)) == ref
    with pytest.raises(ZkbError):
        ctx.prove_end(pkh, h)
    ctx.close()
