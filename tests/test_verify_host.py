"""`B200.verify` — the host-side Groth16 verifier of the product (zokrates_b200/verify.py) on the committed golden proofs:
accepts them, rejects tampered proofs / inputs, agrees with the oracle's pairing, raises on malformed input.  CPU only."""
import copy
import hashlib
import json
import os

import pytest

from oracle import ark, ir as oir
from oracle.ff import BLS12_381, BN254
from zokrates_b200 import backend, curves, proof as pproof

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "proofs.json")
CURVES = {"bn128": BN254, "bls12_381": BLS12_381}


def _case(v):
    c = CURVES[v["curve"]]
    prog = oir.Prog([tuple(a) for a in v["arguments"]], v["return_count"],
                    [oir.Constraint(*[[tuple(t) for t in lc] for lc in cons]) for cons in v["constraints"]])
    w = oir.execute(c, prog, [int(x) for x in v["inputs"]])
    r1cs, z = ark.synthesize(prog, w)
    pk = ark.setup(c, r1cs, ark.Trapdoor(*v["trapdoor"]))
    pk_bytes = ark.pk_serialize(c, pk)
    assert hashlib.sha256(pk_bytes).hexdigest() == v["pk_sha256"]
    vk = pproof.vk_from_pk_bytes(curves.curve(v["curve"]), pk_bytes)
    return c, pk, vk, pproof.Proof.from_json(v["proof_json"]), prog.public_inputs_values(w)


@pytest.mark.parametrize("idx", range(4))
def test_accepts_golden_and_rejects_tampered(idx):
    v = json.load(open(GOLD))[idx]
    c, pk, vk, proof, pub = _case(v)
    assert backend.B200.verify(vk, proof) is True
    bad = copy.deepcopy(proof)
    bad.proof.c, bad.proof.a = proof.proof.a, proof.proof.c            # valid curve points, wrong proof
    assert backend.B200.verify(vk, bad) is False
    if proof.inputs:
        bad = copy.deepcopy(proof)
        bad.inputs[0] = "0x" + (int(proof.inputs[0], 16) ^ 1).to_bytes(32, "big").hex()
        assert backend.B200.verify(vk, bad) is False
        short = copy.deepcopy(proof); short.inputs = short.inputs[:-1]
        with pytest.raises(ValueError, match="MalformedVerifyingKey"):
            backend.B200.verify(vk, short)
    off = copy.deepcopy(proof)
    off.proof.a.y = "0x" + ((int(proof.proof.a.y, 16) + 1) % c.p).to_bytes(c.fq_bytes, "big").hex()
    with pytest.raises(ValueError, match="not on the curve"):
        backend.B200.verify(vk, off)


def test_agrees_with_the_oracle_pairing():
    v = json.load(open(GOLD))[1]                                         # bn128, program with a return value
    c, pk, vk, proof, pub = _case(v)
    raw = proof.to_raw()
    n = c.fq_bytes
    f = [int.from_bytes(raw[i * n:(i + 1) * n], "little") for i in range(8)]
    oproof = ((f[0], f[1]), ((f[2], f[3]), (f[4], f[5])), (f[6], f[7]))
    assert ark.verify(c, pk, pub, oproof) is True and backend.B200.verify(vk, proof) is True
    assert vk.to_tagged_json().startswith('{\n  "scheme": "g16",\n  "curve": "bn128"')


def test_degenerate_g2_point_is_a_failed_verification_not_a_crash(monkeypatch):
    """a vertical line in the Miller loop (G2 input outside the prime-order subgroup) used to surface as a
    TypeError; the reference's verify returns false.  R + (-R) raises DegeneratePoint and verify_proof maps it to False."""
    from zokrates_b200 import verify as V
    v = json.load(open(GOLD))[0]
    c, pk, vk, proof, pub = _case(v)
    pr = V._pairing(v["curve"])
    q = pr.twist(V._g2(proof.proof.b, c.p))
    with pytest.raises(V.DegeneratePoint):
        pr._add(q, (q[0], pr.F.neg(q[1])))

    def degenerate(self, pairs):
        raise V.DegeneratePoint("forced")
    monkeypatch.setattr(V._Pairing, "product_is_one", degenerate)
    assert backend.B200.verify(vk, proof) is False
