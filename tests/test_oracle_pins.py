"""Pins the oracle against the reference's own known-answer tests, and the C
restatement against the python one.  CPU only."""
import hashlib
import random

import numpy as np
import pytest

from oracle import ark, ir
from oracle.ff import BLS12_381, BN254, g1_group, g2_group, pairing_product_is_one
from zokrates_b200._lib import fr_array, fr_from_array

R = BN254.r


def F(v):
    return int(v) % R


class TestFieldKats:
    """zokrates_field/src/bn128.rs:44-241 (values copied as test vectors, not code)."""

    def test_addition(self):
        assert F(65416358 + 68135) == 65484493
        assert F(5 + F(-2)) == 3
        assert F(65416358 + F(-68135)) == 65348223

    def test_subtraction(self):
        assert F(65416358 - 68135) == 65348223
        assert F(65416358 - F(-68135)) == 65484493
        assert F(68135 - 65416358) == 21888242871839275222246405745257275088548364400416034343698204186575743147394

    def test_multiplication(self):
        assert F(32 * 421) == 13472
        assert F(54 * F(-8912)) == 21888242871839275222246405745257275088548364400416034343698204186575808014369
        assert F(F(-54) * F(-12)) == 648
        a = 21888242871839225222246405785257275088694311157297823662689037894645225727
        b = 218882428715392752222464057432572755886923
        assert F(a * b) == 6042471409729479866150380306128222617399890671095126975526159292198160466142

    def test_division_and_pow(self):
        assert F(48 * pow(12, -1, R)) == 4
        res = F(F(-54) * pow(12, -1, R))
        assert F(12 * res) == F(-54)
        assert pow(54, 12, R) == 614787626176508399616

    def test_required_bits(self):
        assert R.bit_length() == 254     # bn128.rs: get_required_bits() == 254
        assert BLS12_381.r.bit_length() == 255

    def test_c_oracle_matches(self, oracle_c):
        random.seed(3)
        for cid, c in ((0, BN254), (1, BLS12_381)):
            for field, mod, nl in ((0, c.r, 4), (1, c.p, c.fq_bytes // 8)):
                a = [0, 1, mod - 1] + [random.randrange(mod) for _ in range(200)]
                b = [mod - 1, 0, mod - 1] + [random.randrange(mod) for _ in range(200)]
                A, B = fr_array(a, nl), fr_array(b, nl)
                assert fr_from_array(oracle_c.field_op(cid, field, 0, A, B)) == [x * y % mod for x, y in zip(a, b)]
                assert fr_from_array(oracle_c.field_op(cid, field, 1, A, B)) == [(x + y) % mod for x, y in zip(a, b)]
                assert fr_from_array(oracle_c.field_op(cid, field, 2, A, B)) == [(x - y) % mod for x, y in zip(a, b)]
                assert fr_from_array(oracle_c.field_op(cid, field, 3, A[2:40], None)) == [pow(x, -1, mod) for x in a[2:40]]


class TestFormats:
    def test_curve_ids(self):
        # zokrates_book/src/toolbox/ir.md: bn128 curve id 0xb4f7b5bd; Field::id() = sha256(modulus LE)[..4]
        assert hashlib.sha256(BN254.r.to_bytes(32, "little")).digest()[:4].hex() == "b4f7b5bd"
        assert hashlib.sha256(BLS12_381.r.to_bytes(32, "little")).digest()[:4].hex() == "40d8c1f9"

    def test_witness_roundtrip_and_json(self):
        # zokrates_ast/src/ir/witness.rs:108-154
        w = {ir.var_new(42): 42, ir.var_public(8): 8, ir.ONE: 1}
        assert ir.witness_read(ir.witness_write(w)) == w
        assert ir.witness_json(w) == '{\n  "~out_8": "8",\n  "~one": "1",\n  "_42": "42"\n}'

    def test_variable_display(self):
        # zokrates_ast/src/common/flat/variable.rs:82-102
        assert ir.var_name(ir.ONE) == "~one"
        assert ir.var_name(ir.var_public(0)) == "~out_0" and ir.var_name(ir.var_public(42)) == "~out_42"
        assert ir.var_name(ir.var_new(0)) == "_0" and ir.var_name(ir.var_new(42)) == "_42"

    def test_sha256_packed_kat_digest(self):
        # zokrates_stdlib/tests/tests/hashes/sha256/512bitPacked.json:5-16 — sha256packed(0,0,0,5)
        d = hashlib.sha256(b"\0" * 63 + b"\x05").digest()
        assert int.from_bytes(d[:16], "big") == 263561599766550617289250058199814760685
        assert int.from_bytes(d[16:], "big") == 65303172752238645975888084098459749904


class TestSolvers:
    """zokrates_interpreter/src/lib.rs:426-511."""

    def test_condition_eq(self):
        assert ir.execute_solver(BN254, "ConditionEq", None, [0]) == [0, 1]
        assert ir.execute_solver(BN254, "ConditionEq", None, [1]) == [1, 1]

    def test_bits(self):
        res = ir.execute_solver(BN254, "Bits", 254, [1])
        assert res[253] == 1 and not any(res[:253])
        res = ir.execute_solver(BN254, "Bits", 254, [42])
        assert res[247:] == [0, 1, 0, 1, 0, 1, 0]
        assert ir.execute_solver(BN254, "Bits", 500, [1]) == [0] * 499 + [1]


class TestRng:
    def test_chacha_structure_against_rfc7539(self):
        # the 20-round variant of the same block function reproduces RFC 7539 §2.3.2's first words
        key = bytes(range(32))
        rng = ark.ChaCha12Rng(key)
        init = [0x61707865, 0x3320646E, 0x79622D32, 0x6B206574] + rng.key + [1, 0x09000000, 0x4A000000, 0]
        s = list(init)
        for _ in range(10):
            for q in ((0, 4, 8, 12), (1, 5, 9, 13), (2, 6, 10, 14), (3, 7, 11, 15), (0, 5, 10, 15), (1, 6, 11, 12),
                      (2, 7, 8, 13), (3, 4, 9, 14)):
                ark.ChaCha12Rng._qr(s, *q)
        out = [(x + y) & 0xFFFFFFFF for x, y in zip(s, init)]
        assert out[:4] == [0xE4E7F110, 0x15593BD1, 0x1FDD0F50, 0xC47120A3]

    def test_chacha12_published_vector(self):
        """ChaCha12, 256-bit all-zero key, zero nonce / counter: the published keystream block (draft-strombergson-chacha-test-
        vectors-01, TC1, 12 rounds — the vector rand_chacha's own ChaCha12 tests use).  Pins the 12-round core itself, oracle and
        product, to something outside this repository; the word order of `next_u32` (little-endian words in state order) with it."""
        want = bytes.fromhex("9bf49a6a0755f953811fce125f2683d50429c3bb49e074147e0089a52eae155f"
                             "0564f879d27ae3c02ce82834acfa8c793a629f2ca0de6919610be82f411326be")
        from zokrates_b200.rng import StdRng
        for rng in (ark.ChaCha12Rng(bytes(32)), StdRng(bytes(32))):
            assert b"".join(rng.next_u32().to_bytes(4, "little") for _ in range(16)) == want

    def test_blake2b_seed_published_vector(self):
        """get_rng_from_entropy hashes the entropy with Blake2b-512 (rng.rs:5-20); hashlib's implementation against the RFC 7693
        appendix A vector ("abc")."""
        import hashlib
        assert hashlib.blake2b(b"abc", digest_size=64).hexdigest().startswith("ba80a53f981c4d0d6a2797b69f12f6e94c212f14685ac4b74b12bb6fdbffa2d1")

    def test_fr_rand_in_range_and_deterministic(self):
        for c in (BN254, BLS12_381):
            a = ark.fr_rand(c, ark.rng_from_entropy("entropy"))
            b = ark.fr_rand(c, ark.rng_from_entropy("entropy"))
            assert a == b and 0 <= a < c.r
            assert ark.fr_rand(c, ark.rng_from_entropy("other")) != a

    def test_product_rng_matches_oracle(self):
        from zokrates_b200 import curves, rng
        for name, oc in (("bn128", BN254), ("bls12_381", BLS12_381)):
            pr = rng.get_rng_from_entropy("hello world")
            orr = ark.rng_from_entropy("hello world")
            assert [pr.next_u64() for _ in range(40)] == [orr.next_u64() for _ in range(40)]
            assert rng.fr_rand(curves.curve(name), rng.get_rng_from_entropy("x")) == ark.fr_rand(oc, ark.rng_from_entropy("x"))


def _factorize_prog():
    a_, b_ = ir.var_new(0), ir.var_new(1)
    return ir.Prog([(a_, True), (b_, False)], 0, [ir.Constraint([(a_, 1)], [(a_, 1)], [(b_, 1)])])


class TestGroth16Oracle:
    @pytest.mark.parametrize("c", [BN254, BLS12_381], ids=lambda c: c.name)
    def test_factorize_config1(self, c):
        """BASELINE config 1: a*a == b with 337, 113569 — ark order z = [1, b, a]; proof verifies and equals
        the trapdoor prediction (independent of the NTT / MSM code)."""
        prog = _factorize_prog()
        w = ir.execute(c, prog, [337, 113569])
        r1cs, z = ark.synthesize(prog, w)
        assert z == [1, 113569, 337] and (r1cs.num_instance, r1cs.num_witness) == (2, 1)
        td = ark.Trapdoor(11, 22, 33, 44, 55555, 3, 7)
        pk = ark.setup(c, r1cs, td)
        pkb = ark.pk_serialize(c, pk)
        assert ark.pk_serialize(c, ark.pk_deserialize(c, pkb)) == pkb
        proof, inputs = ark.generate_proof(c, prog, w, pkb, ark.rng_from_entropy("e"))
        rng = ark.rng_from_entropy("e")
        r, s = ark.fr_rand(c, rng), ark.fr_rand(c, rng)
        assert proof == ark.trapdoor_expected_proof(c, r1cs, td, z, r, s)
        assert inputs == [113569]
        assert ark.verify(c, pk, inputs, proof)
        bad = (proof[0], proof[1], g1_group(c).add(proof[2], c.g1))
        assert not ark.verify(c, pk, inputs, bad)

    def test_unsatisfied_constraint(self):
        with pytest.raises(ir.UnsatisfiedConstraint):
            ir.execute(BN254, ir.Prog([(1, True), (2, False)], 0,
                                      [ir.Constraint([(1, 1)], [(1, 1)], [(2, 1)]), ir.Constraint([(1, 1)], [(0, 1)], [(2, 1)])]),
                       [3, 9])

    def test_pairing_bilinear(self):
        for c in (BN254, BLS12_381):
            G1, G2 = g1_group(c), g2_group(c)
            a, b = 123457, 987651
            assert pairing_product_is_one(c, [(G1.mul(c.g1, a), G2.mul(c.g2, b)), (G1.neg(G1.mul(c.g1, a * b)), c.g2)])

    def test_c_oracle_ntt_msm_prove(self, oracle_c):
        from zokrates_b200 import r1cs as pr1cs
        random.seed(9)
        for cid, c in ((0, BN254), (1, BLS12_381)):
            d = ark.Domain(c, 64)
            x = [random.randrange(c.r) for _ in range(64)]
            assert fr_from_array(oracle_c.ntt(cid, fr_array(x))) == d.fft(x)
            assert fr_from_array(oracle_c.ntt(cid, fr_array(x), True, False)) == d.ifft(x)
            assert fr_from_array(oracle_c.ntt(cid, fr_array(x), False, True)) == d.coset_fft(x)
            assert fr_from_array(oracle_c.ntt(cid, fr_array(x), True, True)) == d.coset_ifft(x)
            G1, G2 = g1_group(c), g2_group(c)
            for n in (0, 1, 7, 40):
                pts = [G1.mul(c.g1, random.randrange(1, c.r)) for _ in range(n)]
                sc = [random.choice([0, 1, c.r - 1, random.randrange(c.r)]) for _ in range(n)]
                if n > 3:
                    pts[2] = None
                got = oracle_c.msm(cid, 1, b"".join(ark.ser_g1(c, p) for p in pts), fr_array(sc), c.fq_bytes)
                assert got == ark.ser_g1(c, G1.msm_naive(pts, sc))
            pts = [G2.mul(c.g2, random.randrange(1, c.r)) for _ in range(9)]
            sc = [random.randrange(c.r) for _ in range(9)]
            assert oracle_c.msm(cid, 2, b"".join(ark.ser_g2(c, p) for p in pts), fr_array(sc), c.fq_bytes) == \
                ark.ser_g2(c, G2.msm_naive(pts, sc))
            # full prove vs trapdoor
            prog = _factorize_prog()
            w = ir.execute(c, prog, [337, 113569])
            r1cs, z = ark.synthesize(prog, w)
            td = ark.Trapdoor(5, 6, 7, 8, 99, 2, 3)
            pkb = ark.pk_serialize(c, ark.setup(c, r1cs, td))
            R1 = pr1cs.R1CS(c.name, 1, 2, 1, *[_csr(m, c.r) for m in (r1cs.a, r1cs.b, r1cs.c)])
            proof, _ = oracle_c.prove(cid, pkb, R1, fr_array(z), 1234, 5678, c.fq_bytes)
            exp = ark.trapdoor_expected_proof(c, r1cs, td, z, 1234, 5678)
            assert proof == ark.ser_g1(c, exp[0]) + ark.ser_g2(c, exp[1]) + ark.ser_g1(c, exp[2])
            # the C trapdoor prediction (closed-form Lagrange coefficients) == the python one (same formula, big ints)
            assert oracle_c.trapdoor_expected(cid, R1, [5, 6, 7, 8, 99, 2, 3], fr_array(z), 1234, 5678, c.fq_bytes) == proof

    def test_c_trapdoor_prediction_mid_size(self, oracle_c):
        """zko_trapdoor_expected against the python prediction and the C prover on synthetic circuits with linear-combination
        rows, a public input and (BLS12-381) the 384-bit base field; tau inside the domain is reported, not mis-evaluated."""
        from zokrates_b200 import synthetic
        for cid, c in ((0, BN254), (1, BLS12_381)):
            r1, z = synthetic.make(c.name, 300)
            td = [3, 5, 7, 11, 1234567, 17, 19]
            got = oracle_c.trapdoor_expected(cid, r1, td, z, 111, 222, c.fq_bytes)
            ref, _ = oracle_c.prove(cid, oracle_c.setup(cid, r1, td), r1, z, 111, 222, c.fq_bytes)
            assert got == ref
            rows = [[(int(col), v) for col, v in zip(cols[int(rp[i]):int(rp[i + 1])], fr_from_array(vals[int(rp[i]):int(rp[i + 1])]))]
                    for rp, cols, vals in r1.matrices() for i in range(r1.num_constraints)]
            N = r1.num_constraints
            o = ark.R1CS(r1.num_instance, r1.num_witness, rows[:N], rows[N:2 * N], rows[2 * N:])
            exp = ark.trapdoor_expected_proof(c, o, ark.Trapdoor(*td), fr_from_array(z), 111, 222)
            assert got == ark.ser_g1(c, exp[0]) + ark.ser_g2(c, exp[1]) + ark.ser_g1(c, exp[2])
            with pytest.raises(RuntimeError, match="rc=3"):
                oracle_c.trapdoor_expected(cid, r1, td[:4] + [1] + td[5:], z, 111, 222, c.fq_bytes)   # tau = 1 = w^0


def _csr(rows, r):
    rowptr = [0]
    col, val = [], []
    for row in rows:
        for cidx, k in row:
            col.append(cidx)
            val.append(k % r)
        rowptr.append(len(col))
    return np.array(rowptr, dtype=np.uint64), np.array(col, dtype=np.uint32), fr_array(val)
