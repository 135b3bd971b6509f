"""GM17: oracle restatement of ark-gm17 (oracle/gm17.py, PARITY UNPINNED against real ark-gm17 output) and
the GPU prover (zkb_gm17_prove) against it.  Checked: both pairing equations, the trapdoor prediction (Fr arithmetic only), byte
equality GPU == oracle, on both curves; -m gpu repeats it through libzkb200.so at sizes where the tiled NTT and real MSMs run."""
import io
import random

import pytest

from oracle import ark, gm17, ir as oir
from oracle.ff import BLS12_381, BN254
from tests.util import proof_bytes, rand_prog_pair
from zokrates_b200 import backend, ir as pir, rng as prng
from zokrates_b200.ir import Interpreter

CURVES = {"bn128": BN254, "bls12_381": BLS12_381}


def _case(curve_name, ncons, npub, npriv, seed):
    c = CURVES[curve_name]
    oprog, pprog, inputs = rand_prog_pair(c, ncons, npub, npriv, seed=seed, curve_name=curve_name)
    ow = oir.execute(c, oprog, inputs)
    r1cs, z = ark.synthesize(oprog, ow)
    return c, oprog, pprog, inputs, ow, r1cs, z


def _check(lib, curve_name, ncons, npub, npriv, seed, verify=True):
    c, oprog, pprog, inputs, ow, r1cs, z = _case(curve_name, ncons, npub, npriv, seed)
    td = gm17.Gm17Trapdoor(3 + seed, 5, 7, 1234567 + seed, 11, 13)
    pk = gm17.setup(c, r1cs, td)
    pk_bytes = gm17.pk_serialize(c, pk)
    # the device setup writes the same key bytes as the oracle's
    assert backend.B200.setup_gm17(pprog, [td.alpha, td.beta, td.gamma, td.tau, td.g1_k, td.g2_k], lib=lib) == pk_bytes
    entropy = f"gm17-{seed}"
    orng = ark.rng_from_entropy(entropy)
    d1, d2, r = ark.fr_rand(c, orng), ark.fr_rand(c, orng), ark.fr_rand(c, orng)
    exp = gm17.trapdoor_expected_proof(c, r1cs, td, z, d1, d2, r)
    pub = oprog.public_inputs_values(ow)
    pw = Interpreter().execute(pprog, inputs)
    proof = backend.B200.generate_proof_gm17(pprog, pw, io.BytesIO(pk_bytes), prng.get_rng_from_entropy(entropy), lib=lib)
    assert proof.to_raw() == proof_bytes(c, exp)              # GPU == trapdoor prediction (canonical affine bytes)
    assert proof.input_values() == pub and proof.scheme == "gm17"
    assert '"scheme": "gm17"' in proof.to_tagged_json()
    if verify:
        assert gm17.verify(c, pk, pub, exp)
        # the product's host verifier (B200.verify for GM17) agrees with the oracle's: accepts the proof, rejects a wrong input
        from zokrates_b200.proof import Proof, gm17_vk_from_pk_bytes
        from zokrates_b200.verify import verify_proof_gm17
        from zokrates_b200.curves import curve as pcurve
        vk = gm17_vk_from_pk_bytes(pcurve(curve_name), pk_bytes)
        assert verify_proof_gm17(vk, proof)
        bad = Proof.from_raw(pcurve(curve_name), proof.to_raw(), [(pub[0] + 1) % c.r] + pub[1:], scheme="gm17")
        assert not verify_proof_gm17(vk, bad)
        assert '"scheme": "gm17"' in vk.to_tagged_json()
    return c, pk, r1cs, z, td


@pytest.mark.parametrize("curve_name", ["bn128", "bls12_381"])
def test_oracle_gm17_self_consistent(curve_name):
    c, oprog, pprog, inputs, ow, r1cs, z = _case(curve_name, 6, 2, 2, seed=1)
    td = gm17.Gm17Trapdoor(3, 5, 7, 11, 13, 17)
    pk = gm17.setup(c, r1cs, td)
    assert gm17.pk_deserialize(c, gm17.pk_serialize(c, pk)) == pk
    rnd = random.Random(9)
    pub = oprog.public_inputs_values(ow)
    for masks in ((0, 0, 0), tuple(rnd.randrange(c.r) for _ in range(3))):
        proof = gm17.prove(c, pk, r1cs, z, *masks)
        assert proof == gm17.trapdoor_expected_proof(c, r1cs, td, z, *masks)
        assert gm17.verify(c, pk, pub, proof)
    G1 = ark.g1_group(c) if hasattr(ark, "g1_group") else None
    from oracle.ff import g1_group
    bad = (proof[0], proof[1], g1_group(c).add(proof[2], c.g1))
    assert not gm17.verify(c, pk, pub, bad)
    assert not gm17.verify(c, pk, [(pub[0] + 1) % c.r] + pub[1:], proof)


@pytest.mark.parametrize("curve_name,shape", [("bn128", (5, 1, 2)), ("bn128", (1, 1, 1)), ("bls12_381", (7, 2, 1))])
def test_gm17_prover_emu(curve_name, shape, emu_lib):
    _check(emu_lib, curve_name, *shape, seed=2)


def test_gm17_errors_emu(emu_lib):
    from zokrates_b200._lib import Context, ZkbError
    c, pk, r1cs, z, td = _check(emu_lib, "bn128", 4, 1, 1, seed=3, verify=False)
    ctx = Context(0, 0, emu_lib)
    data = gm17.pk_serialize(c, pk)
    for bad in (data[:100], data[:-1], data + b"\x00"):
        with pytest.raises(ZkbError):
            ctx.gm17_pk_load(bad)
    h = ctx.gm17_pk_load(data)
    ctx.gm17_pk_free(h)
    with pytest.raises(ZkbError):
        ctx.gm17_pk_free(h)


@pytest.mark.gpu
@pytest.mark.parametrize("curve_name,ncons", [("bn128", 40), ("bls12_381", 40), ("bn128", 300)])
def test_gm17_prover_gpu(curve_name, ncons, gpu_lib):
    """300 constraints -> SAP domain 2^10: the tiled NTT passes and bucket MSMs with thousands of points."""
    _check(gpu_lib, curve_name, ncons, 2, 3, seed=4, verify=ncons < 100)
