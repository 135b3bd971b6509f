"""BASELINE config 2 (sha256packed preimage, ~all-bits witness): the in-house circuit is pinned to the
reference's KAT and proven through the C ABI; the proof equals the ark-equivalent CPU prover's bytes and
satisfies the Groth16 pairing equation."""
import numpy as np
import pytest

from oracle import ark
from oracle.ff import BN254
from zokrates_b200 import sha256_circuit
from zokrates_b200._lib import Context, fr_from_array

KAT = (263561599766550617289250058199814760685, 65303172752238645975888084098459749904)


@pytest.fixture(scope="module")
def circuit():
    return sha256_circuit.make("bn128", (0, 0, 0, 5))


def test_kat_and_satisfaction(circuit):
    # zokrates_stdlib/tests/tests/hashes/sha256/512bitPacked.json:5-16
    r1, z, outs = circuit
    assert outs == KAT
    zi = fr_from_array(z)
    assert zi[0] == 1 and tuple(zi[1:3]) == KAT and zi[3:7] == [0, 0, 0, 5]
    R = BN254.r

    def rows(mat):
        rp, cl, vl = mat
        vv = fr_from_array(vl)
        return [sum(vv[k] * zi[int(cl[k])] for k in range(int(rp[i]), int(rp[i + 1]))) % R for i in range(len(rp) - 1)]
    A, B, C = rows(r1.a), rows(r1.b), rows(r1.c)
    assert all(a * b % R == c for a, b, c in zip(A, B, C))
    assert 50000 < r1.num_constraints < 60000 and r1.domain_size == 1 << 16
    # another input: different digest, still satisfied at the outputs
    _, z2, outs2 = sha256_circuit.make("bn128", (1, 2, 3, 4))
    assert outs2 != KAT and tuple(fr_from_array(z2[1:3])) == outs2


@pytest.mark.gpu
def test_prove_sha256_on_gpu(circuit, gpu_lib, oracle_c):
    r1, z, outs = circuit
    ctx = Context(0, 0, gpu_lib)
    h = ctx.r1cs_load(r1.num_constraints, r1.num_instance, r1.num_witness, r1.matrices())
    td = [0x1111, 0x2222, 0x3333, 0x4444, 0x123456789, 3, 7]
    pk = ctx.setup(h, td)
    pkh = ctx.pk_load(pk)
    proof = ctx.prove(pkh, h, z, 1234567, 7654321)
    ref, _ = oracle_c.prove(0, pk, r1, z, 1234567, 7654321, 32)
    assert proof == ref
    # pairing check against the vk inside the key, public inputs = the two digest halves
    c = BN254
    A, off = ark.de_g1(c, proof, 0)
    B, off = ark.de_g2(c, proof, off)
    Cc, off = ark.de_g1(c, proof, off)
    # parse only the vk prefix of the key
    alpha, o = ark.de_g1(c, pk, 0)
    beta, o = ark.de_g2(c, pk, o)
    gamma, o = ark.de_g2(c, pk, o)
    delta, o = ark.de_g2(c, pk, o)
    n = int.from_bytes(pk[o:o + 8], "little"); o += 8
    abc = []
    for _ in range(n):
        p, o = ark.de_g1(c, pk, o)
        abc.append(p)
    key = ark.ProvingKey(alpha, beta, gamma, delta, abc, None, None, [], [], [], [], [])
    assert ark.verify(c, key, list(outs), (A, B, Cc))
    assert not ark.verify(c, key, [outs[0], outs[1] + 1], (A, B, Cc))
    ctx.close()
